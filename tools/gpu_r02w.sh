#!/bin/bash
# Round 2, pass w (1 GPU): grid-feature producer: step eager / graphed with and without tile plans, ncu launch list by kernel
mkdir -p gpurun_out
NERAF_MEGA_PLAN=static timeout 300 python tools/gridnet_quick.py 128 bf16 --graph > gpurun_out/gridnet_quick_static.log 2>&1; echo "static rc=$?"; cp gpurun_out/gridnet_128_bf16.json gpurun_out/gridnet_128_bf16_static.json
NERAF_MEGA_PLAN_VERBOSE=1 timeout 300 python tools/gridnet_quick.py 128 bf16 --graph > gpurun_out/gridnet_quick_auto.log 2>&1; echo "auto rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/gridnet_128_bf16_static.json', 'gpurun_out/gridnet_128_bf16.json'):
    d = json.load(open(f)); print(f, {k: d.get(k) for k in ('ms_per_step', 'launches_per_step')}, (d.get('graphed') or {}).get('ms_per_step'))
PY
grep "mega plan" gpurun_out/gridnet_quick_auto.log | sort | uniq -c | sort -rn | head -40
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/gridnet_launches.csv python tools/gridnet_quick.py 128 bf16 > gpurun_out/gridnet_ncu.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.DictReader([l for l in open('gpurun_out/gridnet_launches.csv') if not l.startswith('==')]))
n = len(rows)
# last step = last n/ (2 warm + 5 timed + eval...) : take launches between the last two 'im2col' of the stem? simpler: aggregate all, divide by steps
acc = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r['Kernel Name'].split('(')[0][-60:]
    acc[k][0] += 1; acc[k][1] += float(r['Metric Value'].replace(',', '')) / 1e3
tot = sum(v[1] for v in acc.values())
print('launches', n, 'total us', round(tot))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{v[1]:10.0f} us {100*v[1]/tot:5.1f}%  {v[0]:5d} x  avg {v[1]/v[0]:7.1f}  {k}")
mg = sorted([float(r['Metric Value'].replace(',', ''))/1e3 for r in rows if 'mega' in r['Kernel Name']], reverse=True)
print('mega launches:', len(mg), 'top', [round(x) for x in mg[:24]])
PY
