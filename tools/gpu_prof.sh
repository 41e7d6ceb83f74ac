#!/bin/bash
# per-layer job timings + ncu full capture of the job-list kernel inside one bench step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_jobs.py > gpurun_out/bench_jobs.log 2>&1; echo "bench_jobs rc=$?"
cat gpurun_out/bench_jobs.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mega -s 9 -c 3 -f -o gpurun_out/mega \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --gl-rirs 0 > gpurun_out/ncu_mega.log 2>&1; echo "ncu rc=$?"
