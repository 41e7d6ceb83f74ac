#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_griffinlim.py tests/test_gpu_field.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python tools/gl_bench.py 2>&1 | tail -4
