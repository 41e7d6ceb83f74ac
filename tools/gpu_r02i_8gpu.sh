#!/bin/bash
# Round 2, pass i (8 GPUs): kernel gradient exchange vs NCCL, exchange timeline, bench line.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
ONLY="nccl bf16,kernel" NERAF_COMM_TRACE=1 timeout 240 $TR tools/time_dp_segments.py > gpurun_out/time_dp8.txt 2>&1; echo "time rc=$?"; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/time_dp8.txt | tail -6 | cut -c1-900
timeout 400 $TR bench.py --gpus 8 --steps 100 --warmup 5 --gl-rirs 0 --grid-net 0 --loss-columns 0 --no-cpu-baseline > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_8gpu.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_8gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')})
except Exception as e: print('no bench line', e)
PY
