#!/bin/bash
# Round 2, pass al (2 GPUs): the backward's last gradient chunk in two halves; one system fence per worker block
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run() { ONLY="kernel" NERAF_COMM_TRACE=1 timeout 200 $TR tools/time_dp_segments.py > gpurun_out/time_dp2_$1.txt 2>&1; echo "$1 rc=$?"; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/time_dp2_$1.txt | grep "kernel  \|rank 0" | cut -c1-220; grep "rank 0" gpurun_out/time_dp2_$1.txt | sed 's/.*| 0.0 MB/| 0.0 MB/' | cut -c1-400; }
NERAF_EXCHANGE_TAIL_MB=64 run tail_whole
run tail_halves
NERAF_COMM_ONE_FENCE=1 run tail_halves_one_fence
timeout 240 $TR tools/check_dp_equals_single.py > gpurun_out/check_dp2.txt 2>&1; echo "check rc=$?"; grep 'kernel exchange\|DP ==' gpurun_out/check_dp2.txt | cut -c1-300
