#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 300 python bench.py --no-cpu-baseline --steps 100 --gl-rirs 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['clocks'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --gl-rirs 0 > gpurun_out/bench_ncu.log 2>&1; echo "ncu list rc=$?"
