#!/bin/bash
# Round 2, pass r (1 GPU): pull-form gradient exchange on a one-rank group: parity test, then graph + parts timing, pull vs push
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_field.py -m gpu -q -x -k "data_parallel" 2>&1 | tail -3
ONE_RANK_GROUP=1 ONLY=kernel timeout 300 python tools/time_dp_parts.py > gpurun_out/dp_parts_pull.txt 2>&1; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/dp_parts_pull.txt | tail -2 | cut -c1-400
ONE_RANK_GROUP=1 ONLY=kernel NERAF_EXCHANGE_PULL=0 timeout 300 python tools/time_dp_parts.py > gpurun_out/dp_parts_push.txt 2>&1; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/dp_parts_push.txt | tail -2 | cut -c1-400
