#!/bin/bash
# Round 2, pass u (1 GPU): ncu --set full of the helper kernels of one step (B = 2048): why they run at 3-4 TB/s
mkdir -p gpurun_out
OFF="--gl-rirs 0 --no-cpu-baseline --large-batch 0 --grid-net 0 --sweep= --no-soundspaces --loss-columns 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'grid_grads|field_prep|pack_list|head_backward|loss_sums' -s 20 -c 5 -f -o gpurun_out/helpers \
  python bench.py --steps 2 --warmup 4 --no-graph $OFF > gpurun_out/ncu_helpers.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/helpers.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/helpers_ncu_summary.txt
ncu -i gpurun_out/helpers.ncu-rep --page details 2>/dev/null | grep -E "^  [a-z_:A-Z]+.*\(|Duration|DRAM Throughput|Memory Throughput|Achieved Occupancy|Theoretical Occupancy|Registers Per|Mem Busy|Max Bandwidth|L2 Cache Throughput|No Eligible|Issued Warp|Stall|Est. Speedup|OPT " | cut -c1-220 > gpurun_out/helpers_details.txt
wc -l gpurun_out/helpers_details.txt
