"""GPU tuning aid: times the Griffin-Lim kernel alone (RAF / SoundSpaces shapes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neraf_b200 import synthetic as syn  # noqa: E402
from neraf_b200.griffinlim import GriffinLim  # noqa: E402

dev = torch.device("cuda:0")
n = int(os.environ.get("GL_N", "1184"))
for shape in (syn.RAF, syn.SOUNDSPACES):
    gl = GriffinLim(n_fft=shape.n_fft, win_length=shape.win, hop_length=shape.hop, power=1)
    g = torch.Generator().manual_seed(0)
    log_d = (torch.randn(n, shape.T, shape.C, shape.F, generator=g) * 1.5 - 3.0).to(dev)
    init = torch.rand(n, shape.C, shape.F, shape.T, dtype=torch.complex64, device=dev)
    for _ in range(2):
        gl.render(log_d, init)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    reps = 3
    for _ in range(reps):
        gl.render(log_d, init)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print(f"{shape.name}: {n} RIRs x {shape.C} ch  {ms:.2f} ms  -> {n / ms * 1e3:.0f} RIR/s  ({n * shape.C / ms * 1e3:.0f} signals/s)")
