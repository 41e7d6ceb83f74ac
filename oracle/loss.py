"""Oracle: spectral loss of the acoustic field and its closed-form gradient (test infrastructure only).

Follows ``STFTLoss`` / ``SpectralConvergenceLoss`` / ``LogSTFTMagnitudeLoss``
(/root/reference/NeRAF/NeRAF_evaluator.py:76-108, :8-26, :29-53) and the weights
applied in ``NeRAFAudioModel.get_loss_dict`` (NeRAF_model.py:584-600):
``audio_sc_loss *= 0.1 * loss_factor``, ``audio_mag_loss *= 1.0 * loss_factor``,
plain-MSE criterion ``audio_mse = mse * loss_factor`` (:594-595).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

EPS_MAG = 1e-3


def stft_loss(x_log: torch.Tensor, y_log: torch.Tensor, loss_type: str = "mse", dtype=torch.float64) -> Dict[str, torch.Tensor]:
    """NeRAF_evaluator.py:101-108 evaluated in ``dtype`` (float64 = the parity yard-stick)."""
    x = x_log.to(dtype)
    y = y_log.to(dtype)
    x_mag = torch.exp(x) - EPS_MAG
    y_mag = torch.exp(y) - EPS_MAG
    sc = torch.sqrt(((y_mag - x_mag) ** 2).sum()) / torch.sqrt((y_mag ** 2).sum())     # :26 Frobenius ratio
    d = y - x
    mag = (d * d).mean() if loss_type == "mse" else d.abs().mean()                       # :50-53
    return {"audio_sc_loss": sc, "audio_mag_loss": mag}


def loss_dict(pred: torch.Tensor, gt: torch.Tensor, criterion: str = "SC+SLMSE", loss_factor: float = 1e-3,
              dtype=torch.float64) -> Dict[str, torch.Tensor]:
    """NeRAF_model.py:584-600."""
    if criterion == "MSE":
        d = pred.to(dtype) - gt.to(dtype)
        return {"audio_mse": (d * d).mean() * loss_factor}
    out = stft_loss(pred, gt, "mse" if "MSE" in criterion else "l1", dtype)
    out["audio_sc_loss"] = out["audio_sc_loss"] * 1e-1 * loss_factor
    out["audio_mag_loss"] = out["audio_mag_loss"] * 1.0 * loss_factor
    return out


def loss_sums(pred: torch.Tensor, gt: torch.Tensor, dtype=torch.float64) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """The four global sums the fused kernel reduces: S_num, S_den, S_sq, S_abs."""
    x = pred.to(dtype)
    y = gt.to(dtype)
    xm = torch.exp(x) - EPS_MAG
    ym = torch.exp(y) - EPS_MAG
    d = y - x
    return ((ym - xm) ** 2).sum(), (ym ** 2).sum(), (d * d).sum(), d.abs().sum()


def loss_grad(pred: torch.Tensor, gt: torch.Tensor, criterion: str = "SC+SLMSE", loss_factor: float = 1e-3,
              g_sc: float = 1.0, g_mag: float = 1.0, dtype=torch.float64) -> torch.Tensor:
    """d(sum of the loss dict, each entry scaled by its upstream gradient)/d pred, closed form.

    d sc/dx_i  = (xm_i - ym_i) e^{x_i} / (sqrt(S_num) sqrt(S_den));
    d mse/dx_i = 2 (x_i - y_i)/N;  d l1/dx_i = sign(x_i - y_i)/N.
    """
    x = pred.to(dtype)
    y = gt.to(dtype)
    n = x.numel()
    if criterion == "MSE":
        return g_mag * loss_factor * 2.0 * (x - y) / n
    s_num, s_den, _, _ = loss_sums(pred, gt, dtype)
    ex = torch.exp(x)
    xm = ex - EPS_MAG
    ym = torch.exp(y) - EPS_MAG
    dsc = (xm - ym) * ex / (torch.sqrt(s_num) * torch.sqrt(s_den))
    if "MSE" in criterion:
        dmag = 2.0 * (x - y) / n
    else:
        dmag = torch.sign(x - y) / n
    return g_sc * 0.1 * loss_factor * dsc + g_mag * loss_factor * dmag
