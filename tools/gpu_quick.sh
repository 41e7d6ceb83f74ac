#!/bin/bash
# quick loop: parity tests + per-layer timings + bench line + per-tile timeline of the forward / backward lists
mkdir -p gpurun_out; rm -f gpurun_out/trace.bin
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_jobs.py 2>&1 | tail -8 > gpurun_out/bench_jobs.log; cat gpurun_out/bench_jobs.log
NERAF_MEGA_TRACE=gpurun_out/trace.bin NCU_SEQ=1 timeout 300 python tools/bench_jobs.py > gpurun_out/trace.log 2>&1; echo "trace rc=$?"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'GL', d.get('griffinlim',{}).get('value'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --gl-rirs 0 > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"
