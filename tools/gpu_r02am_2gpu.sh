#!/bin/bash
# Round 2, pass am (2 GPUs): final data-parallel code: DP == single check, GPU tests of the exchange, bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 240 $TR tools/check_dp_equals_single.py > gpurun_out/check_dp2.txt 2>&1; echo "check rc=$?"; grep 'exchange\|DP ==' gpurun_out/check_dp2.txt | cut -c1-260
timeout 300 python -m pytest tests/test_gpu_field.py -m gpu -q -x -k "data_parallel" 2>&1 | tail -2
timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 --gl-rirs 0 --grid-net 0 --loss-columns 0 --no-cpu-baseline --sweep= --no-soundspaces > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d.get('dp_equals_single'))"
