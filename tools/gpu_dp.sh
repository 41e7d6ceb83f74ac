#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_field.py -m gpu -q -x 2>&1 | tail -2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_dp_nvls.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3 | cut -c1-120; echo "check rc=${PIPESTATUS[0]}"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/time_dp_segments.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --gl-rirs 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], (d.get('e2e_graphed') or {}).get('value'))"
