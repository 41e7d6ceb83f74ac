#!/bin/bash
# Round 2, pass ad (2 GPUs): finer row groups of the gradient exchange
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for mb in 7 5 3; do
  ONLY="kernel" NERAF_EXCHANGE_CHUNK_MB=$mb timeout 200 $TR tools/time_dp_segments.py > gpurun_out/time_dp2_chunk$mb.txt 2>&1; echo "chunk $mb MB rc=$?"; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/time_dp2_chunk$mb.txt | grep "kernel  " | cut -c1-200
done
