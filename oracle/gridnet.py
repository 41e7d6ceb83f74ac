"""Oracle: the grid-feature producer (test infrastructure only -- nothing under neraf_b200/ imports this).

A functional restatement of ``ResNet3D`` (/root/reference/NeRAF/NeRAF_resnet3d.py:116-201) on a reference-layout
``state_dict`` with torch's own convolution / batch-norm / pooling on the CPU, any dtype (the parity tests run it in
float64); gradients come from autograd.  Stem :119-122,179-182; ``Bottleneck.forward`` :95-113; ``BasicBlock.forward``
:57-77; shortcut ``downsample`` :171-175; average pooling window :141-157; stages :184-191.

PINNED: tests/golden/gridnet_resnet50.npz was produced by the reference's real ``ResNet3D_helper`` (importable without
any stub) by oracle/make_golden_gridnet.py; tests/test_gridnet.py checks this restatement against it.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

EPS, MOMENTUM = 1e-5, 0.1       # nn.BatchNorm3d defaults


def avgpool_window(grid_step: Optional[float], n_features: int) -> int:
    """NeRAF_resnet3d.py:138-157."""
    if grid_step is None:
        grid_step = 1 / 128
    if grid_step >= 1 / 64 - 1 / 512:
        return 2 if n_features == 2048 else 4
    if grid_step >= 1 / 128 - 1 / 512:
        return 4 if n_features == 2048 else 8
    return 8 if n_features == 2048 else 16


def _bn(sd, name, x, training, new_stats):
    rm, rv = sd[name + ".running_mean"].to(x.dtype).clone(), sd[name + ".running_var"].to(x.dtype).clone()
    y = F.batch_norm(x, rm, rv, sd[name + ".weight"], sd[name + ".bias"], training, MOMENTUM, EPS)
    if training and new_stats is not None:
        new_stats[name + ".running_mean"], new_stats[name + ".running_var"] = rm, rv
    return y


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, grid_step: Optional[float] = None, n_features: int = 1024,
            training: bool = True, prefix: str = "backbone_net.",
            new_stats: Optional[Dict[str, torch.Tensor]] = None, gates=None, gate_log=None) -> torch.Tensor:
    """x (1, C, D, H, W) -> (1, N, d, h, w) exactly as ``ResNet3D.forward``; ``sd`` values must already have x's dtype
    (parameters may require grad).  ``new_stats`` receives the running statistics a training-mode pass leaves.

    The gradient of a ReLU network is discontinuous in its forward pass: a pre-activation within rounding of zero gives
    a different gate -- and different gradients everywhere behind it -- in float32 than in float64.  To compare a
    lower-precision implementation's BACKWARD with this oracle, the gates can be pinned: ``gate_log`` (a list) receives
    the boolean pattern ``z > 0`` of every ReLU in call order; ``gates`` (such a list) replaces ``relu(z)`` by
    ``z * gates[i]``."""
    p = prefix
    n_relu = [0]

    def relu(z):
        i = n_relu[0]
        n_relu[0] += 1
        if gate_log is not None:
            gate_log.append(z.detach() > 0)
        if gates is None:
            return F.relu(z)
        return z * gates[i].to(z.dtype)

    h = F.conv3d(x, sd[p + "conv1.weight"], stride=2, padding=2)
    h = relu(_bn(sd, p + "bn1", h, training, new_stats))
    h = F.max_pool3d(h, kernel_size=3, stride=2, padding=1)
    n_stages = 4 if n_features == 2048 else 3
    for s in range(n_stages):
        b = 0
        while f"{p}layer{s + 1}.{b}.conv1.weight" in sd:
            q = f"{p}layer{s + 1}.{b}."
            stride = 2 if (s > 0 and b == 0) else 1
            res = h
            if q + "conv3.weight" in sd:                                               # Bottleneck
                o = relu(_bn(sd, q + "bn1", F.conv3d(h, sd[q + "conv1.weight"]), training, new_stats))
                o = relu(_bn(sd, q + "bn2", F.conv3d(o, sd[q + "conv2.weight"], stride=stride, padding=1), training, new_stats))
                o = _bn(sd, q + "bn3", F.conv3d(o, sd[q + "conv3.weight"]), training, new_stats)
            else:                                                                      # BasicBlock
                o = relu(_bn(sd, q + "bn1", F.conv3d(h, sd[q + "conv1.weight"], stride=stride, padding=1), training, new_stats))
                o = _bn(sd, q + "bn2", F.conv3d(o, sd[q + "conv2.weight"], padding=1), training, new_stats)
            if q + "downsample.0.weight" in sd:
                res = _bn(sd, q + "downsample.1", F.conv3d(h, sd[q + "downsample.0.weight"], stride=stride), training, new_stats)
            h = relu(o + res)
            b += 1
    return F.avg_pool3d(h, avgpool_window(grid_step, n_features), stride=1)


def forward_backward(sd: Dict[str, torch.Tensor], x: torch.Tensor, dout: torch.Tensor, grid_step=None, n_features=1024,
                     training=True, dtype=torch.float64, prefix="backbone_net.", gates=None, gate_log=None
                     ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """Feature, gradients of every parameter for the upstream gradient ``dout``, and the updated running statistics
    (``gates`` / ``gate_log``: see ``forward``)."""
    work = {}
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            t = v.detach().to(dtype).clone()
            if not (k.endswith("running_mean") or k.endswith("running_var")):
                t.requires_grad_(True)
            work[k] = t
        else:
            work[k] = v
    stats: Dict[str, torch.Tensor] = {}
    out = forward(work, x.to(dtype), grid_step, n_features, training, prefix, stats, gates, gate_log)
    out.backward(dout.to(dtype).reshape(out.shape))
    grads = {k: v.grad for k, v in work.items() if isinstance(v, torch.Tensor) and v.requires_grad}
    return out.detach(), grads, stats
