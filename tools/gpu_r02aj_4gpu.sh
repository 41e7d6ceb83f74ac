#!/bin/bash
# Round 2, pass ab (8 GPUs): the bench line at N = 8 (RAF value, SoundSpaces step, batch sweep, DP == single check inside)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517"
( time timeout 600 $TR bench.py --gpus 4 --steps 100 --warmup 5 --gl-rirs 0 --grid-net 0 --loss-columns 0 --no-cpu-baseline > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -2 gpurun_out/bench_4gpu.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_4gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d.get('dp_equals_single'), 'ss', {k:(d.get('soundspaces') or {}).get(k) for k in ('value','ms_per_step')})
    print('sweep', [(p['global_batch'], round(p['ms_per_step'],3), round(p['roofline_frac_per_gpu'],3)) for p in d['batch_sweep']['points']])
except Exception as e: print('no bench line', e)
PY
