#!/bin/bash
# One gpurun call for the grid-feature producer (ResNet3D-50): its GPU tests (the xfail-marked ones run for real), the
# 128^3 step eager and as one CUDA graph, the per-operator breakdown, the same with the plain kernel forms (A/B), and
# an ncu launch list of one step.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gridnet_gpu.py -m gpu --runxfail -q > gpurun_out/gridnet_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gridnet_gpu.log
tail -5 gpurun_out/gridnet_gpu.log
timeout 300 python tools/gridnet_quick.py 128 bf16 --graph > gpurun_out/gridnet_quick.log 2>&1; echo "quick rc=$?"; tail -1 gpurun_out/gridnet_quick.log
timeout 300 python tools/gridnet_breakdown.py 128 bf16 > gpurun_out/gridnet_breakdown.log 2>&1; echo "breakdown rc=$?"; tail -20 gpurun_out/gridnet_breakdown.log
NERAF_GRID_SCALAR=1 timeout 300 python tools/gridnet_quick.py 128 bf16 > gpurun_out/gridnet_quick_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/gridnet_quick_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/gridnet_launches.csv python tools/gridnet_quick.py 128 bf16 > gpurun_out/gridnet_ncu.log 2>&1; echo "ncu list rc=$?"
# one full capture each of the producer's HBM-bound passes (largest instances: skip the warm-up launches)
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:im2col_vec8|col2im_vec8|bn_backward_reduce|bn_apply_vec8|column_sums" -s 40 -c 10 -f -o gpurun_out/gridnet_kernels python tools/gridnet_quick.py 128 bf16 > gpurun_out/gridnet_ncu_full.log 2>&1; echo "ncu full rc=$?"
