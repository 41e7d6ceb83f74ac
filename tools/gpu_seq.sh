#!/bin/bash
mkdir -p gpurun_out
NCU_SEQ=1 timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none -k regex:mega --csv --log-file gpurun_out/seq_warm.csv python tools/bench_jobs.py > gpurun_out/seq.log 2>&1; echo "rc=$?"
