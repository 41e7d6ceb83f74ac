#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/bench.err | cut -c1-400
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
print('sweep', [(p['global_batch'], round(p['ms_per_step'],3), round(p['roofline_frac_per_gpu'],3)) for p in d['batch_sweep']['points']])
print('soundspaces', d.get('soundspaces'))
print('loss', d['spectral_loss']['forward'], d['spectral_loss']['backward'])
print('grid', {k: d['grid_feature'].get(k) for k in ('ms_per_step','graphed','cpu_baseline')})
print('gl', d['griffinlim']['value'], d['griffinlim']['roofline']['frac'], 'cpu', d['cpu_baseline'])
PY
