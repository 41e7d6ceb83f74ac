#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'frac', d['roofline']['frac'], 'GL', d.get('griffinlim',{}).get('value'), d['clocks'])
for k in ('e2e','e2e_graphed_no_prefetch','e2e_resident_feed','e2e_eager'):
    print(k, d[k]['value'], d[k]['ms_per_step'])
print('metrics', d.get('acoustic_metrics'))
print('render', d.get('render',{}).get('value'))
print('cpu', d.get('cpu_baseline'))
PY
