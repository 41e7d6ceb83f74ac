#!/bin/bash
# 2-GPU data-parallel run of bench.py exactly as the driver launches it
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_dp2.json 2> gpurun_out/bench_dp2.err; echo "dp2 rc=$?"
tail -3 gpurun_out/bench_dp2.err; cat gpurun_out/bench_dp2.json | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --no-graph > gpurun_out/bench_dp2_eager.json 2> gpurun_out/bench_dp2_eager.err; echo "dp2 eager rc=$?"
tail -3 gpurun_out/bench_dp2_eager.err; cat gpurun_out/bench_dp2_eager.json | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>&1 | tail -2 | cut -c1-400
