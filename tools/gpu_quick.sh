#!/bin/bash
# quick loop: parity tests + per-layer timings + bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_jobs.py > gpurun_out/bench_jobs.log 2>&1; echo "bench_jobs rc=$?"
cat gpurun_out/bench_jobs.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'GL', d.get('griffinlim',{}).get('value'))
PY
