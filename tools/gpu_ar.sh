#!/bin/bash
N=$(nvidia-smi -L | wc -l)
NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/time_allreduce.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -14
