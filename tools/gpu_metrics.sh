#!/bin/bash
timeout 600 python -m pytest tests/test_metrics.py tests/test_gpu_render.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python - <<'PY'
import torch, time
from neraf_b200 import synthetic as syn
from neraf_b200.metrics import acoustic_metrics
for shape, adv in ((syn.RAF, True), (syn.SOUNDSPACES, False)):
    n = 2072
    w = torch.randn(n, shape.C, shape.hop*(shape.T-1), device="cuda") * torch.exp(-torch.arange(shape.hop*(shape.T-1), device="cuda")/4000.0)
    for _ in range(2): acoustic_metrics(w, shape.fs, adv)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(3): acoustic_metrics(w, shape.fs, adv)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/3
    print(shape.name, "metrics ms", dt*1e3, "RIR/s", n/dt)
PY
