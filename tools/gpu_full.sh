#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --shape SoundSpaces --no-cpu-baseline --steps 50 > gpurun_out/bench_ss.json 2> gpurun_out/bench_ss.err; echo "bench SS rc=$?"; tail -2 gpurun_out/bench_ss.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"
python - <<'PY'
import json
for f in ('bench','bench_ss','bench_ref'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    print(f, {k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d.get('roofline',{}).get('frac'),
          'GL', d.get('griffinlim',{}).get('value'), d.get('griffinlim',{}).get('roofline',{}).get('frac'), 'render', d.get('render',{}).get('value'),
          'cpu', d.get('cpu_baseline',{}).get('value'), d.get('griffinlim',{}).get('cpu_baseline',{}).get('value'), d.get('clocks'))
PY
