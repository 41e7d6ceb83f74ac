#!/bin/bash
# Round 2, pass d: first run of the fused loss kernel, programmatic dependent launch, one-rank kernel exchange.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python tools/ab_step.py > gpurun_out/ab_pdl.txt 2>&1; echo "ab rc=$?"; cat gpurun_out/ab_pdl.txt
NERAF_PDL=0 timeout 300 python tools/ab_step.py > gpurun_out/ab_nopdl.txt 2>&1; echo "ab(no pdl) rc=$?"; cat gpurun_out/ab_nopdl.txt
timeout 600 python bench.py --grid-net 0 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'large', d.get('large_batch'), 'loss', d.get('spectral_loss'))
PY
