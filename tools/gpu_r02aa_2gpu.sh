#!/bin/bash
# Round 2, pass aa (2 GPUs): the bench line at N = 2 (DP check inside), exchange timeline
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 --gl-rirs 0 --grid-net 0 --loss-columns 0 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -2 gpurun_out/bench_2gpu.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d.get('dp_check'), 'ss', (d.get('soundspaces') or {}).get('ms_per_step'))
    print('sweep', [(p['global_batch'], round(p['ms_per_step'],3), round(p['roofline_frac_per_gpu'],3)) for p in d['batch_sweep']['points']])
except Exception as e: print('no bench line', e)
PY
ONLY="nccl bf16,kernel" NERAF_COMM_TRACE=1 timeout 240 $TR tools/time_dp_segments.py > gpurun_out/time_dp2.txt 2>&1; echo "time rc=$?"; grep -v 'Warning\|OMP\|\*\*\*' gpurun_out/time_dp2.txt | tail -4 | cut -c1-700
