#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_field.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/time_dp_segments.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -7
