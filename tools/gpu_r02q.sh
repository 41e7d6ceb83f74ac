#!/bin/bash
# Round 2, pass q (1 GPU): K2 with the FMA-pipe exponential -- parity, timing at 65 536 columns, ncu --set full of both kernels
mkdir -p gpurun_out

OFF="--gl-rirs 0 --no-cpu-baseline --large-batch 0 --grid-net 0 --sweep= --no-soundspaces"
timeout 300 python bench.py --steps 20 --warmup 3 $OFF > gpurun_out/bench_k2.json 2> gpurun_out/bench_k2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_k2.json').read().strip().splitlines()[-1])
print('step', d['ms_per_step'], 'loss', d['spectral_loss']['forward'], d['spectral_loss']['backward'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'loss_sums_kernel|loss_backward_kernel' -c 6 -f -o gpurun_out/k2 \
  python bench.py --steps 1 --warmup 1 $OFF > gpurun_out/ncu_k2.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/k2.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/k2_ncu_summary.txt; tail -60 gpurun_out/k2_ncu_summary.txt | cut -c1-160
