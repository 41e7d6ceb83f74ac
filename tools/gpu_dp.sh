#!/bin/bash
for cfg in "32 36" "24 28"; do
set -- $cfg
echo "NCCL_MAX_CTAS=$1 NERAF_COMM_CTAS=$2"
NCCL_MAX_CTAS=$1 NERAF_COMM_CTAS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/time_dp_segments.py 2>&1 | grep -v "^\*\|OMP_NUM\|NCCL version" | tail -4
done
