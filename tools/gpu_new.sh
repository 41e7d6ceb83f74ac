#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_datafeed.py tests/test_gpu_render.py -m gpu -q -x 2>&1 | tail -8
timeout 600 python bench.py --steps 100 --no-cpu-baseline --gl-rirs 0 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
for k in ('e2e','e2e_graphed_no_prefetch','e2e_resident_feed','e2e_eager'):
    print(k, d[k]['value'], d[k]['ms_per_step'])
PY
