#!/bin/bash
mkdir -p gpurun_out
for cfg in "NERAF_PDL=0 NERAF_FUSED_LOSS=0" "NERAF_PDL=0 NERAF_FUSED_LOSS=0 NERAF_GRID_SCALAR=1"; do
  echo "== $cfg"; env $cfg timeout 300 python tools/debug_step_variants.py bf16 2>&1 | grep -v 'Warning\|ld = \|return Variable\|Consider' | tail -24
done > gpurun_out/debug_variants.txt 2>&1
head -90 gpurun_out/debug_variants.txt
