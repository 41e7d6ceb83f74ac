"""Oracle: Griffin-Lim as configured by NeRAF (test infrastructure only).

Reference call sites: ``GriffinLim(n_fft=(N_freq-1)*2, win_length, hop_length, power=1)``
(/root/reference/NeRAF/NeRAF_model.py:139), used at :229 and :753-754 on
``clip(exp(stft) - 1e-3, 0, 1e4)`` (:746-747).  The algorithm lives in
torchaudio (2.1.2 pinned by README.md:49; 2.11 installed here):
``torchaudio.functional.griffinlim`` (functional.py:255-353) on top of
``torch.stft`` / ``torch.istft``.  This file restates all three explicitly
(framing, centred zero-padded window, reflect padding, overlap-add, window-envelope
division, trimming) with only the DFT itself delegated to ``torch.fft``; it is
pinned bit-for-bit / to rounding against ``torchaudio.transforms.GriffinLim`` by
tests/test_oracle.py and the golden vectors of oracle/make_golden.py.

The start phase is an explicit argument (``torch.rand(..., dtype=cfloat)`` in the
reference: Re, Im ~ U[0,1), NOT unit modulus) so both sides can share it.
"""
from __future__ import annotations

from typing import Optional

import torch


def padded_window(n_fft: int, win_length: int, dtype=torch.float32) -> torch.Tensor:
    """hann(win_length, periodic) centred and zero-padded to n_fft (torch.stft/istft semantics)."""
    w = torch.hann_window(win_length, periodic=True, dtype=dtype)
    left = (n_fft - win_length) // 2
    out = torch.zeros(n_fft, dtype=dtype)
    out[left:left + win_length] = w
    return out


def window_envelope(n_fft: int, win_length: int, hop: int, n_frames: int, dtype=torch.float32) -> torch.Tensor:
    """Overlap-added window**2 over the un-trimmed length n_fft + hop*(T-1) (torch.istft)."""
    w2 = padded_window(n_fft, win_length, dtype) ** 2
    env = torch.zeros(n_fft + hop * (n_frames - 1), dtype=dtype)
    for t in range(n_frames):
        env[t * hop:t * hop + n_fft] += w2
    return env


def istft(spec: torch.Tensor, n_fft: int, hop: int, win_length: int) -> torch.Tensor:
    """torch.istft(center=True, normalized=False, onesided=True, length=None). spec: (N, F, T) complex."""
    rdtype = spec.real.dtype
    n, f, t = spec.shape
    w = padded_window(n_fft, win_length, rdtype)
    frames = torch.fft.irfft(spec, n=n_fft, dim=1) * w[None, :, None]          # (N, n_fft, T)
    full = torch.zeros(n, n_fft + hop * (t - 1), dtype=rdtype)
    for i in range(t):
        full[:, i * hop:i * hop + n_fft] += frames[:, :, i]
    env = window_envelope(n_fft, win_length, hop, t, rdtype)
    start = n_fft // 2
    end = full.shape[1] - n_fft // 2
    return full[:, start:end] / env[start:end]


def stft(x: torch.Tensor, n_fft: int, hop: int, win_length: int) -> torch.Tensor:
    """torch.stft(center=True, pad_mode='reflect', normalized=False, onesided=True). x: (N, L) -> (N, F, T)."""
    pad = n_fft // 2
    xp = torch.nn.functional.pad(x[:, None, :], (pad, pad), mode="reflect")[:, 0, :]
    w = padded_window(n_fft, win_length, x.dtype)
    frames = xp.unfold(1, n_fft, hop)                                            # (N, T, n_fft)
    return torch.fft.rfft(frames * w, dim=2).transpose(1, 2)


def griffinlim(mag: torch.Tensor, init_phase: Optional[torch.Tensor], n_fft: int, hop: int, win_length: int,
               n_iter: int = 32, momentum: float = 0.99, power: float = 1.0) -> torch.Tensor:
    """torchaudio.functional.griffinlim (functional.py:255-353). mag: (..., F, T) -> (..., hop*(T-1)).

    ``init_phase=None`` is ``rand_init=False`` (all-ones start).
    """
    m = momentum / (1 + momentum)
    shape = mag.shape
    spec = mag.reshape(-1, shape[-2], shape[-1]).pow(1 / power)
    cdtype = torch.complex64 if spec.dtype == torch.float32 else torch.complex128
    if init_phase is None:
        angles = torch.ones(spec.shape, dtype=cdtype)
    else:
        angles = init_phase.reshape(spec.shape).to(cdtype)
    tprev = torch.zeros((), dtype=cdtype)
    for _ in range(n_iter):
        inverse = istft(spec * angles, n_fft, hop, win_length)
        rebuilt = stft(inverse, n_fft, hop, win_length)
        angles = rebuilt - m * tprev
        angles = angles / (angles.abs() + 1e-16)
        tprev = rebuilt
    wave = istft(spec * angles, n_fft, hop, win_length)
    return wave.reshape(shape[:-2] + wave.shape[-1:])


def log_to_mag(log_stft: torch.Tensor) -> torch.Tensor:
    """NeRAF_model.py:746-747: clip(exp(x) - 1e-3, 0, 1e4)."""
    return torch.clip(torch.exp(log_stft) - 1e-3, 0.0, 10000.0)
