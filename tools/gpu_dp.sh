#!/bin/bash
# 2-GPU data-parallel run of bench.py exactly as the driver launches it
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_field.py -m gpu -q -x 2>&1 | tail -3
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_dp2.json 2> gpurun_out/bench_dp2.err; echo "dp2 rc=$?"
tail -3 gpurun_out/bench_dp2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_dp2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], 'GL', d.get('griffinlim',{}).get('value'), d['config']['launch'])
PY
