#!/bin/bash
# Round 2, second GPU call: -m gpu suite (new boundary + producer-in-graph tests), then ncu --set full captures with source
# of the job-list kernel and of the step's helper kernels, and of the stand-alone spectral-loss kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mega -s 6 -c 2 -f -o gpurun_out/mega python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --gl-rirs 0 --grid-net 0 --large-batch 0 --loss-columns 0 > gpurun_out/ncu_mega.log 2>&1; echo "ncu mega rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:pack_list|field_prep|loss_sums|head_backward|grid_grads" -s 15 -c 5 -f -o gpurun_out/helpers python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --gl-rirs 0 --grid-net 0 --large-batch 0 --loss-columns 0 > gpurun_out/ncu_helpers.log 2>&1; echo "ncu helpers rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:loss_sums|loss_backward" -s 420 -c 4 -f -o gpurun_out/loss python bench.py --steps 1 --warmup 3 --no-cpu-baseline --gl-rirs 0 --grid-net 0 --large-batch 0 > gpurun_out/ncu_loss.log 2>&1; echo "ncu loss rc=$?"
