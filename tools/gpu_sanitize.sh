#!/bin/bash
# compute-sanitizer memcheck over the small-size parity tests (every kernel of the library is launched at least once;
# CUDA-graph capture is excluded: the sanitizer's own stream use invalidates captures)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 --print-limit 20 \
  python -m pytest tests/test_datafeed.py tests/test_metrics.py tests/test_gpu_render.py tests/test_gpu_field.py tests/test_gpu_griffinlim.py tests/test_gpu_ops.py tests/test_zz_gridnet_gpu.py -m gpu -q \
  -k "not full_size and not training_size and not many_signals and not acoustic_metrics_within and not graph and not dims4 and not train_step_through" > gpurun_out/sanitize.log 2>&1
echo "sanitizer rc=$?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitize.log
grep -E "ERROR SUMMARY|passed|failed|Invalid|========= (Invalid|Out)" gpurun_out/sanitize.log | head -20
