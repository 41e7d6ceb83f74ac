"""Oracle: the acoustic MLP and its hand-derived backward (test infrastructure only).

Follows ``NeRAFAudioSoundField`` (/root/reference/NeRAF/NeRAF_field.py:37-65):
five trunk ``nn.Linear`` layers each followed by ``leaky_relu(0.1)`` (:49-51),
``sound_rez`` heads ``10 * tanh(Linear(feat))`` (:56-58) stacked on dim 1
(:60-63).  Weights are taken as a reference-layout ``state_dict``
(``soundfield.{i}.{weight,bias}``, ``STFT_linear.{c}.{weight,bias}``) so a real
reference module and this restatement can share parameters.

``field_forward`` is the dense restatement (exactly the reference's operation
sequence, any dtype).  ``field_forward_factored`` hoists the batch-invariant grid
feature out of layer 1 -- the algebra the CUDA path
uses -- and ``field_backward`` is the closed-form gradient checked against
autograd in tests/test_oracle.py.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

NEG_SLOPE = 0.1
N_TRUNK = 5


def split_state(sd: Dict[str, torch.Tensor]) -> Tuple[List[torch.Tensor], List[torch.Tensor], List[torch.Tensor], List[torch.Tensor]]:
    tw = [sd[f"soundfield.{i}.weight"] for i in range(N_TRUNK)]
    tb = [sd[f"soundfield.{i}.bias"] for i in range(N_TRUNK)]
    c = 0
    hw, hb = [], []
    while f"STFT_linear.{c}.weight" in sd:
        hw.append(sd[f"STFT_linear.{c}.weight"])
        hb.append(sd[f"STFT_linear.{c}.bias"])
        c += 1
    return tw, tb, hw, hb


def field_forward(sd: Dict[str, torch.Tensor], h: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """NeRAF_field.py:47-65 verbatim semantics; returns (B, C, F)."""
    tw, tb, hw, hb = split_state(sd)
    x = h.to(dtype)
    for w, b in zip(tw, tb):
        x = F.leaky_relu(F.linear(x, w.to(dtype), b.to(dtype)), negative_slope=NEG_SLOPE)
    outs = [(torch.tanh(F.linear(x, w.to(dtype), b.to(dtype))) * 10).unsqueeze(1) for w, b in zip(hw, hb)]
    return torch.cat(outs, dim=1)


def field_forward_factored(sd, enc: torch.Tensor, grid_feature: torch.Tensor, dtype=torch.float32,
                           keep: bool = False):
    """Layer 1 as ``enc @ W1[:, G:]^T + (b1 + W1[:, :G] @ g)``; identical result up to rounding.

    With ``keep=True`` also returns the activations needed by ``field_backward``.
    """
    tw, tb, hw, hb = split_state(sd)
    g = grid_feature.flatten().to(dtype)
    G = g.numel()
    w1 = tw[0].to(dtype)
    c1 = tb[0].to(dtype) + w1[:, :G] @ g
    acts = [enc.to(dtype)]
    z = F.linear(acts[0], w1[:, G:], c1)
    x = F.leaky_relu(z, NEG_SLOPE)
    acts.append(x)
    for w, b in zip(tw[1:], tb[1:]):
        x = F.leaky_relu(F.linear(x, w.to(dtype), b.to(dtype)), NEG_SLOPE)
        acts.append(x)
    wh = torch.cat([w.to(dtype) for w in hw], dim=0)
    bh = torch.cat([b.to(dtype) for b in hb], dim=0)
    y = 10 * torch.tanh(F.linear(x, wh, bh))
    out = y.view(y.shape[0], len(hw), -1)
    return (out, acts) if keep else out


def field_backward(sd, acts: List[torch.Tensor], grid_feature: torch.Tensor, out: torch.Tensor,
                   dout: torch.Tensor, dtype=torch.float64):
    """Closed-form backward of the factored field (autograd of NeRAF_field.py:47-65 +
    the ``expand``/``cat`` at NeRAF_model.py:557-560).

    Returns (grads: dict in state_dict naming, dgrid: (G,)).
    dZ_head = dout * (10 - out^2/10); dZ_l = (dZ_{l+1} W_{l+1}) * leaky'(x_l);
    dW_l = dZ_l^T x_{l-1}; db_l = colsum dZ_l;
    dW_1[:, :G] = db_1 (outer) g; dgrid = W_1[:, :G]^T db_1.
    """
    tw, tb, hw, hb = split_state(sd)
    C = len(hw)
    B = out.shape[0]
    y = out.reshape(B, -1).to(dtype)
    dy = dout.reshape(B, -1).to(dtype)
    dz = dy * (10.0 - y * y / 10.0)
    grads = {}
    Fq = hw[0].shape[0]
    x5 = acts[5].to(dtype)
    dwh = dz.t() @ x5
    dbh = dz.sum(0)
    for c in range(C):
        grads[f"STFT_linear.{c}.weight"] = dwh[c * Fq:(c + 1) * Fq]
        grads[f"STFT_linear.{c}.bias"] = dbh[c * Fq:(c + 1) * Fq]
    wh = torch.cat([w.to(dtype) for w in hw], dim=0)
    dx = dz @ wh
    g = grid_feature.flatten().to(dtype)
    G = g.numel()
    dgrid = None
    for l in range(N_TRUNK - 1, -1, -1):
        x_l = acts[l + 1].to(dtype)
        dz = dx * torch.where(x_l > 0, torch.ones_like(x_l), torch.full_like(x_l, NEG_SLOPE))
        db = dz.sum(0)
        grads[f"soundfield.{l}.bias"] = db
        if l > 0:
            grads[f"soundfield.{l}.weight"] = dz.t() @ acts[l].to(dtype)
            dx = dz @ tw[l].to(dtype)
        else:
            dw_enc = dz.t() @ acts[0].to(dtype)
            dw_grid = torch.outer(db, g)
            grads["soundfield.0.weight"] = torch.cat([dw_grid, dw_enc], dim=1)
            dgrid = tw[0].to(dtype)[:, :G].t() @ db
    return grads, dgrid
