#!/bin/bash
# batch sweep of the train step on one GPU (BASELINE config 5), plus tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep.jsonl
for b in 1024 2048 4096 8192 16384 32768 65536; do
  timeout 300 python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline --gl-rirs 0 >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    d=json.loads(l)
    print(d['config']['batch_per_gpu'], round(d['value']/1e6,3),'Mcol/s', round(d['ms_per_step'],3),'ms', 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']/1e6,3), d['clocks'])
PY
