#!/bin/bash
# Round 2, pass ae (1 GPU): ncu --set full of the widening + grid-gradient launch and of the exchange kernel (one-rank group)
mkdir -p gpurun_out
ONE_RANK_GROUP=1 ONLY=kernel N_ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'grid_grads|loss_head' -s 8 -c 4 -f -o gpurun_out/dpk \
  python tools/time_dp_parts.py > gpurun_out/ncu_dpk.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/dpk.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/dpk_summary.txt
ncu -i gpurun_out/dpk.ncu-rep --page details 2>/dev/null | grep -E "grid_grads_kernel|loss_head_kernel|Duration|DRAM Throughput|L2 Cache Throughput|Mem Busy|Max Bandwidth|Achieved Occupancy|Theoretical Occupancy|Issued Warp|No Eligible|Stall|stall|L1/TEX Hit|L2 Hit|Mem Pipes Busy|OPT|Est\." | cut -c1-260 | head -120
