#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/trace16k.bin
COMMON="--steps 1 --warmup 3 --no-graph --gl-rirs 0 --no-cpu-baseline --large-batch 0 --loss-columns 0 --grid-net 0 --sweep= --no-soundspaces"
NERAF_MEGA_TRACE=gpurun_out/trace16k.bin timeout 300 python bench.py --batch 16384 $COMMON > gpurun_out/trace16k.log 2>&1; echo "rc=$?"
python - <<'PY'
import sys
sys.path.insert(0, 'tools')
import mega_trace as mt
f = 'gpurun_out/trace16k.bin'
L = mt.read(f)
sys.argv = ['x', f, str(len(L) - 2), str(len(L) - 1)]
mt.main()
PY
rm -f gpurun_out/trace16k.bin
