#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_griffinlim.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/gl_bench.py 2>&1 | tail -3
GL_N=296 timeout 600 ncu --set full --import-source on --clock-control none -k regex:griffinlim_kernel -s 2 -c 1 -f -o gpurun_out/gl python tools/gl_bench.py > gpurun_out/gl_ncu.log 2>&1; echo "ncu rc=$?"
