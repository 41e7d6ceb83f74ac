#!/bin/bash
# Round 2, pass l (1 GPU): full GPU test suite on HEAD, then the structural cost of the kernel-exchange step on a ONE-rank
# group (no link traffic): graph + eager parts, and an ncu launch list of it.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
ONE_RANK_GROUP=1 timeout 300 python tools/time_dp_parts.py > gpurun_out/dp_parts_1rank.txt 2>&1; echo "parts rc=$?"; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/dp_parts_1rank.txt | tail -4 | cut -c1-400
ONE_RANK_GROUP=1 NERAF_EXCHANGE_CHUNK_MB=11 ONLY=kernel timeout 300 python tools/time_dp_parts.py > gpurun_out/dp_parts_1rank_11mb.txt 2>&1; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/dp_parts_1rank_11mb.txt | tail -2 | cut -c1-400
ONE_RANK_GROUP=1 ONLY=kernel N_ITERS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_dp1.csv python tools/time_dp_parts.py > gpurun_out/ncu_dp1.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows = list(csv.DictReader([l for l in open('gpurun_out/launches_dp1.csv') if not l.startswith('==')]))
# last 14 launches = last eager step
for r in rows[-16:]:
    print(f"{float(r['Metric Value'].replace(',',''))/1e3:8.1f} us  grid {r['Grid Size']:>14} blk {r['Block Size']:>12}  {r['Kernel Name'][:80]}")
PY
