"""Generates tests/golden/gridnet_resnet50.npz from the REFERENCE's own ResNet3D (build container only: needs
/root/reference).  NeRAF_resnet3d.py imports nothing but torch, so the real ``ResNet3D_helper`` runs unmodified:

    python -m oracle.make_golden_gridnet

Inputs are the seeded synthetic grid / state_dict of neraf_b200/synthetic.py (regenerated identically on the GPU box), so
only outputs are stored: the feature in training and evaluation mode, the running statistics a training step leaves,
and -- in evaluation mode, where the gradient is well conditioned -- the gradients of every
batch-norm parameter and of a few convolutions in full plus the norm and a fixed random projection of every gradient.
Also the reference's state_dict keys and shapes (the checkpoint contract).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

N, GRID_STEP = 64, 1 / 64
FULL_CONVS = ("backbone_net.conv1.weight", "backbone_net.layer1.0.conv2.weight", "backbone_net.layer2.0.downsample.0.weight",
              "backbone_net.layer3.5.conv3.weight")


def stable_seed(name: str) -> int:
    return sum((i + 1) * ord(ch) for i, ch in enumerate(name)) % (2 ** 31)


def projection(t: torch.Tensor, name: str) -> float:
    g = torch.Generator().manual_seed(stable_seed(name))
    return float((t.double() * torch.randn(t.shape, generator=g).double()).sum())


def main() -> None:
    from NeRAF.NeRAF_resnet3d import ResNet3D_helper
    from neraf_b200 import synthetic as syn
    torch.manual_seed(0)
    sd = syn.make_gridnet_state_dict("resnet50")
    x = syn.make_grid(N)
    dout = torch.randn(1, 1024, 1, 1, 1, generator=torch.Generator().manual_seed(5))
    out = {}
    ref = ResNet3D_helper(in_channels=7, backbone="resnet50", pretrained=False, grid_step=GRID_STEP, N_features=1024)
    out["state_keys"] = np.array(json.dumps({k: list(v.shape) for k, v in ref.state_dict().items()}))
    ref.load_state_dict(sd, strict=True)

    ref.train()
    with torch.no_grad():
        out["feature_train"] = ref(x).reshape(-1).numpy()
    after = ref.state_dict()
    for k in ("backbone_net.bn1", "backbone_net.layer2.0.downsample.1", "backbone_net.layer3.5.bn3"):
        out["stats/" + k + ".running_mean"] = after[k + ".running_mean"].numpy().copy()
        out["stats/" + k + ".running_var"] = after[k + ".running_var"].numpy().copy()
    out["num_batches_tracked"] = np.array(int(after["backbone_net.bn1.num_batches_tracked"]))

    ref.load_state_dict(sd, strict=True)
    ref.eval()
    y = ref(x)
    out["feature_eval"] = y.detach().reshape(-1).numpy()
    y.backward(dout)
    norms, projs = {}, {}
    for k, p in ref.named_parameters():
        norms[k] = float(p.grad.double().norm())
        projs[k] = projection(p.grad, k)
        if p.dim() == 1 or k in FULL_CONVS:
            out["grad_eval/" + k] = p.grad.numpy().copy()
    out["grad_eval_norms"] = np.array(json.dumps(norms))
    out["grad_eval_projections"] = np.array(json.dumps(projs))
    out["dout"] = dout.reshape(-1).numpy()
    path = os.path.join(ROOT, "tests", "golden", "gridnet_resnet50.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
