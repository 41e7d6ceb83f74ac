#!/bin/bash
# data-parallel record: equality check + bench line on all visible GPUs
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/check_dp_equals_single.py 2>&1 | grep -v "^\*\|OMP_NUM\|NCCL version" | tail -4 | tee gpurun_out/check_dp$N.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err; echo "dp$N rc=$?"
tail -2 gpurun_out/bench_dp$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_dp$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'feed', (d.get('e2e_resident_feed') or {}).get('value'), 'GL', d.get('griffinlim',{}).get('value'), 'render', d.get('render',{}).get('value'))
PY
