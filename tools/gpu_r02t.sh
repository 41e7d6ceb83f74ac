#!/bin/bash
# Round 2, pass t (1 GPU): widening pass with four loads in flight (one-rank group): test, graph + parts, ncu cold durations
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_field.py -m gpu -q -x -k "data_parallel" 2>&1 | tail -2
ONE_RANK_GROUP=1 timeout 300 python tools/time_dp_parts.py > gpurun_out/dp_parts_1rank.txt 2>&1; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/dp_parts_1rank.txt | tail -2 | cut -c1-400
ONE_RANK_GROUP=1 ONLY=kernel N_ITERS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_dp1.csv python tools/time_dp_parts.py > gpurun_out/ncu_dp1.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows = list(csv.DictReader([l for l in open('gpurun_out/launches_dp1.csv') if not l.startswith('==')]))
for r in rows[-8:]:
    print(f"{float(r['Metric Value'].replace(',',''))/1e3:8.1f} us  grid {r['Grid Size']:>14} blk {r['Block Size']:>12}  {r['Kernel Name'][:80]}")
PY
