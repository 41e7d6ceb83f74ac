#!/bin/bash
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 \
  python -m pytest tests/test_gpu_griffinlim.py tests/test_metrics.py tests/test_datafeed.py -m gpu -q -k "other_stft_geometries or zero_iterations or edge_cases or batched_metrics" > gpurun_out/racecheck.log 2>&1
echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|hazard" gpurun_out/racecheck.log | head -12
