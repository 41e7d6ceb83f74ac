#!/bin/bash
# Round 2, pass e: tuning of the loss launch / PDL trigger, ncu launch list of both loss forms.
mkdir -p gpurun_out
for cfg in "NERAF_PDL=0" "NERAF_PDL_TRIGGER=late" "NERAF_PDL_TRIGGER=early" "NERAF_PDL_TRIGGER=late NERAF_LOSS_BLOCKS=6" "NERAF_PDL=0 NERAF_LOSS_BLOCKS=6" "NERAF_PDL=0 NERAF_LOSS_BLOCKS=2"; do
  echo "== $cfg"; env $cfg timeout 300 python tools/ab_step.py 2>&1 | tail -3
done | tee gpurun_out/ab_tuning.txt
NERAF_PDL=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ab.csv python tools/ab_step.py > gpurun_out/ab_ncu.log 2>&1; echo "ncu rc=$?"
for k in 60 61 62; do python tools/launch_list.py gpurun_out/launches_ab.csv $k; done 2>&1 | tail -40
timeout 600 python -m pytest tests/test_zz_gridnet_gpu.py tests/test_gpu_field.py -m gpu -q -x 2>&1 | tail -5
