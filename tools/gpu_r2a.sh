#!/bin/bash
# Round 2, first GPU call: the whole -m gpu suite on the current tree (xfail marks ignored), the producer's step timing
# and per-operator breakdown, the per-tile timeline of the job-list kernel, a bench line.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu --runxfail -q -rA --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -20
timeout 300 python tools/gridnet_quick.py 128 bf16 --graph > gpurun_out/gridnet_quick.log 2>&1; echo "quick rc=$?"; tail -1 gpurun_out/gridnet_quick.log
timeout 300 python tools/gridnet_breakdown.py 128 bf16 > gpurun_out/gridnet_breakdown.log 2>&1; echo "breakdown rc=$?"; tail -30 gpurun_out/gridnet_breakdown.log
rm -f gpurun_out/trace.bin
NERAF_MEGA_TRACE=gpurun_out/trace.bin timeout 300 python bench.py --steps 1 --warmup 3 --no-graph --gl-rirs 0 --no-cpu-baseline --grid-net 0 > gpurun_out/trace.log 2>&1; echo "trace rc=$?"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/gridnet_launches.csv python tools/gridnet_quick.py 128 bf16 > gpurun_out/gridnet_ncu.log 2>&1; echo "ncu list rc=$?"
