#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/trace.bin
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 50 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'GL', d.get('griffinlim',{}).get('value'), d['clocks'])
PY
NERAF_MEGA_TRACE=gpurun_out/trace.bin timeout 300 python bench.py --steps 1 --warmup 3 --no-graph --gl-rirs 0 --no-cpu-baseline > gpurun_out/trace.log 2>&1; echo "trace rc=$?"
