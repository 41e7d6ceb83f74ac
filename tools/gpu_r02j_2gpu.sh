#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_field.py -m gpu -q -x -k "data_parallel" 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
ONLY="nccl bf16,kernel" NERAF_COMM_TRACE=1 timeout 240 $TR tools/time_dp_segments.py > gpurun_out/time_dp2.txt 2>&1; echo "time rc=$?"; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/time_dp2.txt | tail -5 | cut -c1-900
ONLY="kernel" NERAF_MEGA_STAGES=6 NERAF_COMM_TRACE=1 timeout 240 $TR tools/time_dp_segments.py > gpurun_out/time_dp2_6st.txt 2>&1; echo "time(6 stages: exchange behind) rc=$?"; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/time_dp2_6st.txt | tail -3 | cut -c1-900
timeout 240 $TR tools/check_dp_equals_single.py > gpurun_out/check_dp2.txt 2>&1; echo "check rc=$?"; grep 'kernel exchange\|DP ==' gpurun_out/check_dp2.txt | cut -c1-400
