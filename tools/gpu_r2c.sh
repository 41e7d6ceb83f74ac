#!/bin/bash
# Round 2, third GPU call: -m gpu suite, operand-stream microbenchmark (TMA ring without MMA: unicast / shared / multicast),
# one large GEMM through the job-list kernel against cuBLAS.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 120 tools/micro/tma_bw 8 64 > gpurun_out/tma_bw_l2.txt 2>&1; echo "tma_bw l2 rc=$?"; cat gpurun_out/tma_bw_l2.txt
timeout 120 tools/micro/tma_bw 256 2 > gpurun_out/tma_bw_hbm.txt 2>&1; echo "tma_bw hbm rc=$?"; cat gpurun_out/tma_bw_hbm.txt
timeout 300 python tools/gemm_peak.py > gpurun_out/gemm_peak.txt 2>&1; echo "gemm_peak rc=$?"; cat gpurun_out/gemm_peak.txt
