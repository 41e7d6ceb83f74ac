"""Spectral loss of the acoustic field on the fused CUDA reduction (K2).

Mirrors ``STFTLoss`` / ``SpectralConvergenceLoss`` / ``LogSTFTMagnitudeLoss``
(/root/reference/NeRAF/NeRAF_evaluator.py:8-108): ``STFTLoss(loss_type)(x_log, y_log)`` returns
``{'audio_sc_loss', 'audio_mag_loss'}`` (unweighted, like the reference; the model applies
``0.1*loss_factor`` / ``loss_factor``, NeRAF_model.py:597-598).

``spectral_loss`` is the functional form with the weights folded into the kernels and an optional
``torch.distributed`` process group: the Frobenius ratio of the SC term is a *global* quantity, so
under data parallelism the four partial sums are all-reduced before the loss and its gradient are
formed -- the result equals the single-GPU loss on the concatenated batch.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib


class _SpectralLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred: torch.Tensor, gt: torch.Tensor, criterion: int, w_sc: float, w_mag: float, group):
        lib = _lib.lib()
        _lib.require_device(pred, "loss input")
        dev = pred.device
        p = pred.detach().contiguous().float()
        g = gt.detach().to(device=dev, non_blocking=True).contiguous().float()
        if p.shape != g.shape:
            raise ValueError(f"pred {tuple(p.shape)} and target {tuple(g.shape)} differ in shape")
        n = p.numel()
        if n == 0:
            raise ValueError("spectral loss of an empty batch is undefined")
        sums = torch.empty(5, dtype=torch.float64, device=dev)     # 4 sums + the fused kernel's completion ticket
        losses = torch.empty(2, dtype=torch.float32, device=dev)
        stream = _lib.stream_ptr(dev)
        n_total = n
        if group is None:                                          # one launch: reduction + finalize
            _lib.check(lib.neraf_spectral_loss_forward(p.data_ptr(), g.data_ptr(), n, criterion, w_sc, w_mag,
                                                       sums.data_ptr(), losses.data_ptr(), stream))
        else:
            import torch.distributed as dist
            _lib.check(lib.neraf_spectral_loss_sums(p.data_ptr(), g.data_ptr(), n, sums.data_ptr(), 0, stream))
            dist.all_reduce(sums[:4], op=dist.ReduceOp.SUM, group=group)
            n_total = n * dist.get_world_size(group)      # equal shards (the DP sampler guarantees it)
            _lib.check(lib.neraf_spectral_loss_finalize(sums.data_ptr(), n_total, criterion, w_sc, w_mag,
                                                        losses.data_ptr(), stream))
        ctx.save_for_backward(p, g, sums)
        ctx.meta = (criterion, w_sc, w_mag, n_total)
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g_sc: torch.Tensor, g_mag: torch.Tensor):
        lib = _lib.lib()
        p, g, sums = ctx.saved_tensors
        criterion, w_sc, w_mag, n_total = ctx.meta
        dev = p.device
        g_sc = g_sc if g_sc.dtype == torch.float32 else g_sc.float()      # 0-d upstream gradients, read on the device
        g_mag = g_mag if g_mag.dtype == torch.float32 else g_mag.float()
        dpred = torch.empty_like(p)
        _lib.check(lib.neraf_spectral_loss_backward(p.data_ptr(), g.data_ptr(), p.numel(), n_total, criterion,
                                                    sums.data_ptr(), g_sc.data_ptr(), g_mag.data_ptr(), w_sc, w_mag,
                                                    dpred.data_ptr(), _lib.stream_ptr(dev)))
        return dpred, None, None, None, None, None


def spectral_loss(pred: torch.Tensor, gt: torch.Tensor, criterion: str = "SC+SLMSE", w_sc: float = 1.0,
                  w_mag: float = 1.0, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(w_sc * SC, w_mag * log-magnitude term) as 0-d fp32 CUDA tensors with autograd.

    For ``criterion="MSE"`` the first value is 0 and the second is ``w_mag * mse``.
    """
    if criterion not in _lib.CRITERIA:
        raise ValueError(f"criterion must be one of {list(_lib.CRITERIA)}")
    return _SpectralLossFn.apply(pred, gt, _lib.CRITERIA[criterion], float(w_sc), float(w_mag), group)


class SpectralConvergenceLoss(nn.Module):
    """||y_mag - x_mag||_F / ||y_mag||_F on LOG inputs (the exp/-1e-3 of STFTLoss is fused in)."""

    def forward(self, x_log: torch.Tensor, y_log: torch.Tensor) -> torch.Tensor:
        return spectral_loss(x_log, y_log, "SC+SLMSE")[0]


class LogSTFTMagnitudeLoss(nn.Module):
    def __init__(self, loss_type: str = "l1"):
        super().__init__()
        if loss_type not in ("l1", "mse"):
            raise ValueError("loss_type must be 'l1' or 'mse'")
        self.loss_type = loss_type

    def forward(self, x_log: torch.Tensor, y_log: torch.Tensor) -> torch.Tensor:
        return spectral_loss(x_log, y_log, "SC+SLMSE" if self.loss_type == "mse" else "SC+SLL1")[1]


class STFTLoss(nn.Module):
    """Reference-compatible module (NeRAF_evaluator.py:76-108): one fused reduction for both terms."""

    def __init__(self, loss_type: str = "l1", group=None):
        super().__init__()
        if loss_type not in ("l1", "mse"):
            raise ValueError("loss_type must be 'l1' or 'mse'")
        self.loss_type = loss_type
        self.group = group

    def forward(self, x_log: torch.Tensor, y_log: torch.Tensor) -> Dict[str, torch.Tensor]:
        sc, mag = spectral_loss(x_log, y_log, "SC+SLMSE" if self.loss_type == "mse" else "SC+SLL1", group=self.group)
        return {"audio_sc_loss": sc, "audio_mag_loss": mag}
