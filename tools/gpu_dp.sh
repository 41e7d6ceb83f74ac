#!/bin/bash
mkdir -p gpurun_out
for dt in fp32 bf16; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --gl-rirs 0 --grad-dtype $dt > gpurun_out/bench_dp2_$dt.json 2> gpurun_out/bench_dp2.err; echo "dp2 $dt rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_dp2_$dt.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'])
PY
done
