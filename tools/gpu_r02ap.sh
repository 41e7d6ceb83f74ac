#!/bin/bash
# Round 2, pass ap (1 GPU): ncu --set full of the forward / backward job-list launches of the FINAL build (feature-set kernels)
mkdir -p gpurun_out
timeout 70 ncu --set full --clock-control none -k regex:mega -s 6 -c 2 -f -o gpurun_out/mega python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --gl-rirs 0 --grid-net 0 --large-batch 0 --loss-columns 0 --sweep= --no-soundspaces > gpurun_out/ncu_mega.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/mega.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py mega > gpurun_out/mega_ncu_summary_variants.txt
rm -f gpurun_out/mega.ncu-rep
grep -E "^==|time_duration|tensor_cycles|dram__bytes_(read|write).sum " gpurun_out/mega_ncu_summary_variants.txt | cut -c1-140
