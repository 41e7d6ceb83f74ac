#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/trace.bin
NERAF_MEGA_TRACE=gpurun_out/trace.bin timeout 300 python bench.py --steps 1 --warmup 3 --no-graph --gl-rirs 0 --no-cpu-baseline > gpurun_out/trace.log 2>&1; echo "rc=$?"
