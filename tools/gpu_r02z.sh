#!/bin/bash
# Round 2, pass z (1 GPU): grid-gradient kernel (looping row groups): parity + cold duration + step; producer tests (128-bit pooling gradient)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_field.py tests/test_gpu_ops.py tests/test_zz_gridnet_gpu.py -m gpu -q -x 2>&1 | tail -2
OFF="--gl-rirs 0 --no-cpu-baseline --large-batch 0 --grid-net 0 --sweep= --no-soundspaces --loss-columns 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_z.csv python bench.py --steps 2 --warmup 3 $OFF > gpurun_out/ncu_z.log 2>&1; echo "ncu rc=$?"
python tools/launch_list.py gpurun_out/launches_z.csv | tail -9
timeout 300 python bench.py --steps 200 --warmup 10 $OFF 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'])"
ONE_RANK_GROUP=1 timeout 300 python tools/time_dp_parts.py > gpurun_out/dp_parts_1rank.txt 2>&1; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/dp_parts_1rank.txt | tail -2 | cut -c1-400
timeout 300 python tools/gridnet_quick.py 128 bf16 --graph 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('producer', d['ms_per_step'], d.get('graph_ms_per_step'))"
