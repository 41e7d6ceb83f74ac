#!/bin/bash
# Round 2, pass ac (8 GPUs): the largest gradient matrix exchanged in 1 / 2 / 3 row groups
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
for mb in 64 11 7; do
  ONLY="kernel" NERAF_EXCHANGE_CHUNK_MB=$mb NERAF_COMM_TRACE=1 timeout 200 $TR tools/time_dp_segments.py > gpurun_out/time_dp8_chunk$mb.txt 2>&1; echo "chunk $mb MB rc=$?"; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/time_dp8_chunk$mb.txt | grep "kernel  \|rank 0" | cut -c1-1000
done
