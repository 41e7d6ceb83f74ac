#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/loudness_map.py > gpurun_out/loudness_map.json 2> gpurun_out/loudness_map.err; echo "loudness rc=$?"; cat gpurun_out/loudness_map.json | cut -c1-500
timeout 600 python bench.py --shape SoundSpaces > gpurun_out/bench_ss.json 2> gpurun_out/bench_ss.err; echo "ss rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_ss.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'GL', d['griffinlim']['value'], d['griffinlim']['roofline']['frac'], 'metrics', d['acoustic_metrics']['value'])
PY
