#!/bin/bash
# Round 2, pass h (2 GPUs): kernel gradient exchange -- equality with single-process training, timing vs NCCL.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_field.py -m gpu -q -x -k "data_parallel or graphed_step" 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/check_dp_equals_single.py > gpurun_out/check_dp2.txt 2>&1; echo "check rc=$?"; grep -v Warning gpurun_out/check_dp2.txt | tail -8 | cut -c1-400
timeout 300 $TR tools/time_dp_segments.py > gpurun_out/time_dp2.txt 2>&1; echo "time rc=$?"; grep -v Warning gpurun_out/time_dp2.txt | tail -6 | cut -c1-300
NERAF_NO_MULTICAST=1 timeout 300 $TR tools/time_dp_segments.py > gpurun_out/time_dp2_nomc.txt 2>&1; echo "time(no multicast) rc=$?"; grep kernel gpurun_out/time_dp2_nomc.txt | tail -2 | cut -c1-300
timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 --gl-rirs 0 --grid-net 0 --loss-columns 0 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_2gpu.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config'].get('launch'))
except Exception as e: print('no bench line', e)
PY
