"""Prints one train step of an ncu launch list (gpurun_out/launches.csv): kernels in order with their device time."""
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
rows = list(csv.DictReader([l for l in open(path) if not l.startswith("==")]))
# a training step starts with the re-pack of the parameters (render / inference calls do not re-pack)
idx = [i for i, r in enumerate(rows) if "pack_list" in r["Kernel Name"]]
k = int(sys.argv[2]) if len(sys.argv) > 2 else -2          # which step: index into the list of re-pack launches
a, b = idx[k], idx[k + 1] if k + 1 != 0 else len(rows)
tot = 0.0
for r in rows[a:b]:
    v = float(r["Metric Value"].replace(",", "")) / 1e3
    tot += v
    print(f"{v:8.1f} us  grid {r['Grid Size']:>14} blk {r['Block Size']:>12}  {r['Kernel Name'][:90]}")
print(f"total {tot:.1f} us over {b - a} launches")
