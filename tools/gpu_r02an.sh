#!/bin/bash
# Round 2, pass an (1 GPU): feature-specialised variants of the job-list kernel: parity, then step time with / without them
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_field.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -2
OFF="--gl-rirs 0 --no-cpu-baseline --large-batch 0 --grid-net 0 --sweep= --no-soundspaces --loss-columns 0"
for v in 1 0; do
  NERAF_MEGA_VARIANTS=$v timeout 300 python bench.py --steps 300 --warmup 10 $OFF 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variants $v: step', round(d['ms_per_step']*1e3,1), 'us  frac', round(d['roofline']['frac'],4))"
  NERAF_MEGA_VARIANTS=$v ONE_RANK_GROUP=1 ONLY=kernel timeout 300 python tools/time_dp_parts.py 2>&1 | grep "kernel  " | cut -c1-200
done
