#!/bin/bash
mkdir -p gpurun_out
COMMON="--steps 2 --warmup 3 --no-graph --gl-rirs 0 --no-cpu-baseline --large-batch 0 --loss-columns 0 --grid-net 0 --sweep= --no-soundspaces"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_16k.csv python bench.py --batch 16384 $COMMON > gpurun_out/ncu16k.log 2>&1; echo "ncu rc=$?"
python tools/launch_list.py gpurun_out/launches_16k.csv
