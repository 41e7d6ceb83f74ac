#!/bin/bash
mkdir -p gpurun_out
NCU_SEQ=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:mega -s 14 -c 1 -f -o gpurun_out/fwd1 python tools/bench_jobs.py > gpurun_out/seq.log 2>&1; echo "rc=$?"
