#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_field.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline --steps 100 --gl-rirs 0 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','host_wall_ms_per_step')}, 'e2e', d['e2e'])
PY
