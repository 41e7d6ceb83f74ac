#!/bin/bash
# Round 2, pass v (1 GPU): helper kernels with more loads in flight: parity tests, cold durations, step time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_field.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -2
OFF="--gl-rirs 0 --no-cpu-baseline --large-batch 0 --grid-net 0 --sweep= --no-soundspaces --loss-columns 0"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_v.csv python bench.py --steps 2 --warmup 3 $OFF > gpurun_out/ncu_v.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows = list(csv.DictReader([l for l in open('gpurun_out/launches_v.csv') if not l.startswith('==')]))
rows = [r for r in rows if r['Metric Name'] == 'gpu__time_duration.sum']
for r in rows[-9:]:
    print(f"{float(r['Metric Value'].replace(',',''))/1e3:8.1f} us  grid {r['Grid Size']:>14} blk {r['Block Size']:>12}  {r['Kernel Name'][:80]}")
PY
timeout 300 python bench.py --steps 200 --warmup 10 $OFF 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'])"
