#!/bin/bash
# Round 2, pass s (2 GPUs): pull-form exchange (peer loads through shared memory, fp32 delivery) vs the push form
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 240 $TR tools/check_dp_equals_single.py > gpurun_out/check_dp2.txt 2>&1; echo "check rc=$?"; grep 'kernel exchange\|DP ==\|Error\|error' gpurun_out/check_dp2.txt | cut -c1-300 | head
ONLY="kernel" NERAF_COMM_TRACE=1 timeout 240 $TR tools/time_dp_segments.py > gpurun_out/time_dp2_pull.txt 2>&1; echo "time rc=$?"; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/time_dp2_pull.txt | tail -3 | cut -c1-900
ONLY="kernel" NERAF_EXCHANGE_PULL=0 NERAF_COMM_TRACE=1 timeout 240 $TR tools/time_dp_segments.py > gpurun_out/time_dp2_push.txt 2>&1; echo "time rc=$?"; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/time_dp2_push.txt | tail -3 | cut -c1-900
ONLY=single,kernel timeout 240 $TR tools/time_dp_parts.py > gpurun_out/dp_parts_2gpu.txt 2>&1; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/dp_parts_2gpu.txt | tail -3 | cut -c1-400
