#!/bin/bash
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_dp_nvls.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3 | cut -c1-120; echo "check rc=${PIPESTATUS[0]}"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/time_dp_segments.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3
