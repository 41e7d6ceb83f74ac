#!/bin/bash
# One gpurun call for the record: parity tests, smoke, bench line, reference arm, ncu launch list, ncu --set full of the
# job-list kernel.  Outputs under gpurun_out/ (copy what should be judged into profiles/).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -4 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"
# launch list of the SAME command the bench line comes from (CUDA graph replays: ncu lists the kernel nodes), then of the eager plugin calls
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --gl-rirs 148 --grid-net 0 > gpurun_out/bench_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_eager.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --gl-rirs 0 --grid-net 0 > gpurun_out/bench_ncu_eager.log 2>&1; echo "ncu eager list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mega -s 6 -c 2 -f -o gpurun_out/mega python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --gl-rirs 0 --grid-net 0 > gpurun_out/ncu_mega.log 2>&1; echo "ncu full rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'e2e_feed', (d.get('e2e_resident_feed') or {}).get('value'), 'e2e_eager', d['e2e_eager']['value'], 'frac', d['roofline']['frac'], 'GL', d['griffinlim']['value'], d['griffinlim']['roofline']['frac'], 'render', d['render']['value'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
PY
# K2 at 65 536 columns alone (tools/k2_only.py): ncu --set full of the two loss kernels
timeout 300 ncu --set full --clock-control none -k regex:'loss_sums_kernel|loss_backward_kernel' -s 4 -c 2 -f -o gpurun_out/k2 python tools/k2_only.py > gpurun_out/ncu_k2.log 2>&1; echo "ncu k2 rc=$?"
ncu -i gpurun_out/k2.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/k2_ncu_summary.txt
ncu -i gpurun_out/mega.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py mega > gpurun_out/mega_ncu_summary.txt
rm -f gpurun_out/k2.ncu-rep gpurun_out/mega.ncu-rep
grep -E "^==|time_duration|dram__bytes_read.sum |dram__bytes_write.sum |dram_throughput" gpurun_out/k2_ncu_summary.txt | cut -c1-150
