"""K2 alone at 65 536 RAF columns (33.6 M elements, inputs larger than L2): three forward + backward passes of the spectral
loss through the plugin's autograd nodes -- the launches `ncu --set full` captures for profiles/ (tools/gpu_round.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neraf_b200.loss import spectral_loss  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
n_cols, F = 65536, 513
pred = (torch.randn(n_cols, 1, F, device=dev, generator=g) * 2).requires_grad_(True)
gt = torch.randn(n_cols, 1, F, device=dev, generator=g) * 2
for _ in range(3):
    pred.grad = None
    sc, mag = spectral_loss(pred, gt, "SC+SLMSE", 1e-4, 1e-3)
    (sc + mag).backward()
torch.cuda.synchronize()
print(float(sc), float(mag))
