#!/bin/bash
# Round 2, pass x (1 GPU): new tests (GradScaler, B = 2048 oracle parity), producer tests + timing after the pooling-gradient change
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_field.py -m gpu -q -x -k "autocast or full_size" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_zz_gridnet_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python tools/gridnet_quick.py 128 bf16 --graph > gpurun_out/gridnet_quick.log 2>&1; echo "quick rc=$?"; tail -2 gpurun_out/gridnet_quick.log | cut -c1-600
