#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_eager.py > gpurun_out/profile_eager.txt 2>&1; echo "rc=$?"
grep -v "^$" gpurun_out/profile_eager.txt | cut -c1-170 | head -60
