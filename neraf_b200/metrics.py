"""Acoustic metrics of rendered impulse responses, batched on the device.

Mirrors /root/reference/NeRAF/NeRAF_helper.py -- ``compute_t60`` (:48-64), ``evaluate_edt`` (:148-161),
``evaluate_clarity`` (:109-122) -- which the evaluator (NeRAF_evaluator.py:131-190) calls once per RIR on numpy copies
of the Griffin-Lim output.  Here the waveforms stay on the device and ONE launch (``neraf_acoustic_metrics``) measures
the whole batch: T60 (pyroomacoustics' Schroeder fit; RAF: after a 200 Hz high-pass, 10 dB span; SoundSpaces: 30 dB),
EDT and C50.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib


def acoustic_metrics(wave: torch.Tensor, fs: float, advanced: bool = False, t60: bool = True, edt: bool = True,
                     c50: bool = True) -> Dict[str, torch.Tensor]:
    """wave: (..., L) float32 device tensor of impulse responses -> {'t60', 'edt', 'c50'}: float64 tensors of shape (...).

    ``advanced`` selects the RAF variant of T60 (``measure_rt60_advance``: 200 Hz high-pass, decay_db=10) instead of
    ``measure_rt60(decay_db=30)``; a failed fit reads -1 like ``compute_t60``'s except branch."""
    if not wave.is_cuda:
        raise _lib.NerafError("acoustic_metrics: the waveforms must live on a CUDA device (there is no CPU path)")
    lead, L = wave.shape[:-1], wave.shape[-1]
    w = wave.to(torch.float32).contiguous().view(-1, L)
    S, dev = w.shape[0], w.device
    out = {k: torch.empty(S, dtype=torch.float64, device=dev) for k, on in (("t60", t60), ("edt", edt), ("c50", c50)) if on}
    p = _lib.MetricParams()
    p.n_samples, p.fs = L, float(fs)
    p.t60_decay_db, p.t60_highpass_hz = (10.0, 200.0) if advanced else (30.0, 0.0)
    ws = torch.empty(S * L if (advanced and t60) else 1, dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().neraf_acoustic_metrics(C.byref(p), w.data_ptr(), S, ws.data_ptr(), ws.numel() * 4,
                                                 _lib.ptr(out.get("t60")), _lib.ptr(out.get("edt")),
                                                 _lib.ptr(out.get("c50")), _lib.stream_ptr(dev)))
    return {k: v.view(lead) for k, v in out.items()}


def _pair(pred: torch.Tensor, gt: torch.Tensor, key: str, fs: float, advanced: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    both = torch.stack([torch.as_tensor(gt), torch.as_tensor(pred)]).to(torch.float32)
    m = acoustic_metrics(both, fs, advanced, t60=key == "t60", edt=key == "edt", c50=key == "c50")[key].cpu().numpy()
    return m[0], m[1]


def compute_t60(true_in: torch.Tensor, gen_in: torch.Tensor, fs: float, advanced: bool = False):
    """NeRAF_helper.py:48-64: (C, L) ground truth and prediction -> (gt, pred) arrays of C T60 values."""
    return _pair(gen_in, true_in, "t60", fs, advanced)


def evaluate_edt(pred_ir: torch.Tensor, gt_ir: torch.Tensor, fs: float):
    """NeRAF_helper.py:148-161 -> (gt, pred)."""
    return _pair(pred_ir, gt_ir, "edt", fs)


def evaluate_clarity(pred_ir: torch.Tensor, gt_ir: torch.Tensor, fs: float):
    """NeRAF_helper.py:109-122 -> (gt, pred)."""
    return _pair(pred_ir, gt_ir, "c50", fs)
