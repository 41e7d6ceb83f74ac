#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/gridnet_launches.csv python tools/gridnet_quick.py 128 bf16 > gpurun_out/gridnet_ncu.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.DictReader([l for l in open('gpurun_out/gridnet_launches.csv') if not l.startswith('==')]))
acc = collections.defaultdict(list)
for r in rows:
    k = r['Kernel Name'].split('(')[0][-48:]
    acc[k].append(float(r['Metric Value'].replace(',', '')) / 1e3)
tot = sum(sum(v) for v in acc.values())
print('launches', len(rows), 'total us', round(tot))
for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1]))[:16]:
    vs = sorted(v, reverse=True)
    print(f"{sum(v):9.0f} us {100*sum(v)/tot:5.1f}%  {len(v):4d} x  avg {sum(v)/len(v):6.1f}  top {[round(x) for x in vs[:6]]}  median {vs[len(vs)//2]:.1f}  {k}")
PY
