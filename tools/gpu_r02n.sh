#!/bin/bash
# Round 2, pass n (1 GPU): explicit tile plans -- parity under forced plans, then plan vs stride timing over the batch sweep
mkdir -p gpurun_out
NERAF_MEGA_PLAN=cp timeout 900 python -m pytest tests/test_gpu_field.py tests/test_gpu_ops.py -m gpu -q -x > gpurun_out/pytest_plan_cp.log 2>&1; echo "pytest (forced critical-path plans) rc=$?"; tail -3 gpurun_out/pytest_plan_cp.log | cut -c1-200
NERAF_MEGA_PLAN=rb timeout 900 python -m pytest tests/test_gpu_field.py -m gpu -q -x > gpurun_out/pytest_plan_rb.log 2>&1; echo "pytest (forced row-block plans) rc=$?"; tail -3 gpurun_out/pytest_plan_rb.log | cut -c1-200
NERAF_MEGA_PLAN_VERBOSE=1 timeout 600 python tools/ab_plan.py > gpurun_out/ab_plan.txt 2>&1; echo "ab rc=$?"; grep "B=\|mega plan" gpurun_out/ab_plan.txt | cut -c1-250
