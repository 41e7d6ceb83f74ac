#!/bin/bash
# Round 2, pass ak (2 GPUs): exchange timeline with 5 MB row groups; poll back-off 200 ns vs 40 ns
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for ns in 200 40; do
  ONLY="kernel" NERAF_COMM_SLEEP_NS=$ns NERAF_COMM_TRACE=1 timeout 200 $TR tools/time_dp_segments.py > gpurun_out/time_dp2_sleep$ns.txt 2>&1; echo "sleep $ns ns rc=$?"; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/time_dp2_sleep$ns.txt | grep "kernel  \|rank 0" | cut -c1-1300
done
