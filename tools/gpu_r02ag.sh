#!/bin/bash
# Round 2, pass ag (1 GPU): compute-sanitizer memcheck over the parity tests (new this round: the exchange kernel on a one-rank
# group, the one-launch loss, planned tile lists, the 128-bit pooling gradient); captures are excluded (the sanitizer's own
# stream use invalidates them) except the one-rank data-parallel test, which is run eagerly by the tools below
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 --print-limit 20 \
  python -m pytest tests/test_datafeed.py tests/test_metrics.py tests/test_gpu_render.py tests/test_gpu_field.py tests/test_gpu_griffinlim.py tests/test_gpu_ops.py tests/test_zz_gridnet_gpu.py -m gpu -q \
  -k "not full_size and not training_size and not many_signals and not acoustic_metrics_within and not graph and not dims4 and not train_step_through and not 5000 and not data_parallel" > gpurun_out/sanitize.log 2>&1
echo "sanitizer rc=$?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitize.log
grep -E "ERROR SUMMARY|passed|failed|========= (Invalid|Out)" gpurun_out/sanitize.log | head -10
# planned tile lists forced for every job list, under memcheck
NERAF_MEGA_PLAN=cp timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 --print-limit 20 \
  python -m pytest tests/test_gpu_field.py -m gpu -q -k "train_step_matches_oracle and bf16" > gpurun_out/sanitize_plan.log 2>&1
echo "sanitizer (forced plans) rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_plan.log | head -4
