#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log | cut -c1-300
for i in 1 2 3; do timeout 300 python -m pytest tests/test_zz_gridnet_gpu.py -m gpu -q -x -k "train_step_through" 2>&1 | tail -2; done
