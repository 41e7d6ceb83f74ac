"""Batched Griffin-Lim on the fused CUDA kernel (K3).

``GriffinLim`` mirrors ``torchaudio.transforms.GriffinLim`` as the reference constructs it
(/root/reference/NeRAF/NeRAF_model.py:139: ``GriffinLim(n_fft=(N_freq-1)*2, win_length, hop_length,
power=1)`` with torchaudio's defaults n_iter=32, momentum=0.99, rand_init=True, hann window) and
is called like it: ``istft_transform(mag)`` with ``mag`` of shape (..., F, T) -> (..., hop*(T-1)).

``render`` is the batched fast path for field outputs: log-magnitude columns laid out
(N, T, C, F) -- what the acoustic field produces for N poses x T time bins -- go straight to
waveforms (N, C, L) with the log->magnitude conversion of NeRAF_model.py:746-747 fused in and no
permute / host round trip.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _lib


_WS_CACHE = {}


def _workspace(nbytes: int, dev: torch.device) -> torch.Tensor:
    """Grow-only scratch per (device, stream): the kernel's staged magnitudes and previous-waveform lines (tens to
    hundreds of MB per batch) are not handed back to the allocator between renders."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _WS_CACHE[key] = ws
    return ws


def _run(params: _lib.GlParams, n_items: int, n_channels: int, spec: torch.Tensor, strides,
         init: Optional[torch.Tensor], init_strides, out: torch.Tensor) -> None:
    lib = _lib.lib()
    dev = spec.device
    ws_b = C.c_size_t()
    _lib.check(lib.neraf_griffinlim_sizes(C.byref(params), n_items * n_channels, C.byref(ws_b)))
    ws = _workspace(max(ws_b.value, 16), dev)
    _lib.check(lib.neraf_griffinlim(C.byref(params), n_items, n_channels, spec.data_ptr(), *strides,
                                    None if init is None else init.data_ptr(), *init_strides,
                                    ws.data_ptr(), ws.numel(), out.data_ptr(), _lib.stream_ptr(dev)))


class GriffinLim(nn.Module):
    def __init__(self, n_fft: int = 400, n_iter: int = 32, win_length: Optional[int] = None,
                 hop_length: Optional[int] = None, power: float = 2.0, momentum: float = 0.99,
                 length: Optional[int] = None, rand_init: bool = True):
        super().__init__()
        if not 0 <= momentum < 1:
            raise ValueError("momentum must be in the range [0, 1). Found: {}".format(momentum))
        self.n_fft = n_fft
        self.n_iter = n_iter
        self.win_length = win_length if win_length is not None else n_fft
        self.hop_length = hop_length if hop_length is not None else self.win_length // 2
        self.power = power
        self.momentum = momentum
        self.length = length
        self.rand_init = rand_init

    def _params(self, n_frames: int, input_is_log: bool) -> _lib.GlParams:
        p = _lib.GlParams()
        p.n_fft, p.win_length, p.hop, p.n_frames = self.n_fft, self.win_length, self.hop_length, n_frames
        p.n_iter, p.momentum, p.input_is_log = self.n_iter, self.momentum, int(input_is_log)
        return p

    def forward(self, specgram: torch.Tensor, init_phase: Optional[torch.Tensor] = None) -> torch.Tensor:
        """specgram (..., F, T) magnitude**power -> waveform (..., hop*(T-1)).

        ``init_phase`` (complex64, same shape) overrides the random start so two implementations can
        share it; by default it is drawn with ``torch.rand(..., dtype=cfloat)`` exactly like torchaudio.
        """
        _lib.require_device(specgram, "specgram")
        shape = specgram.shape
        F_, T = shape[-2], shape[-1]
        if F_ != self.n_fft // 2 + 1:
            raise ValueError(f"expected {self.n_fft // 2 + 1} frequency bins, got {F_}")
        spec = specgram.reshape(-1, F_, T).float()
        if self.power != 1:
            spec = spec.pow(1.0 / self.power)
        spec = spec.contiguous()
        S = spec.shape[0]
        if init_phase is None and self.rand_init:
            init_phase = torch.rand(spec.shape, dtype=torch.complex64, device=spec.device)
        init_r = None
        if init_phase is not None:
            init_r = torch.view_as_real(init_phase.reshape(S, F_, T).to(torch.complex64).contiguous())
        L = self.hop_length * (T - 1)
        wave = torch.empty(S, L, dtype=torch.float32, device=spec.device)
        _run(self._params(T, False), S, 1, spec, (F_ * T, 0, 1, T), init_r, (F_ * T, 0, 1, T), wave)
        if self.length is not None:                     # torch.istft(length=...) trims / zero-pads at the end
            if self.length <= L:
                wave = wave[:, :self.length]
            else:
                wave = torch.nn.functional.pad(wave, (0, self.length - L))
        return wave.reshape(shape[:-2] + wave.shape[-1:])

    def render(self, log_stft: torch.Tensor, init_phase: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Field output (N, T, C, F) log-magnitudes -> waveforms (N, C, hop*(T-1)).

        ``init_phase``: complex64 (N, C, F, T) (torchaudio layout) or None (random / ones per rand_init).
        """
        _lib.require_device(log_stft, "log_stft")
        if log_stft.dim() != 4:
            raise ValueError("expected (N, T, C, F)")
        x = log_stft.float().contiguous()
        N, T, Cc, F_ = x.shape
        if F_ != self.n_fft // 2 + 1:
            raise ValueError(f"expected {self.n_fft // 2 + 1} frequency bins, got {F_}")
        if init_phase is None and self.rand_init:
            init_phase = torch.rand(N, Cc, F_, T, dtype=torch.complex64, device=x.device)
        L = self.hop_length * (T - 1)
        wave = torch.empty(N, Cc, L, dtype=torch.float32, device=x.device)
        init_r = None if init_phase is None else torch.view_as_real(init_phase.to(torch.complex64).contiguous())
        _run(self._params(T, True), N, Cc, x, (T * Cc * F_, F_, Cc * F_, 1), init_r,
             (Cc * F_ * T, F_ * T, 1, T), wave)
        return wave
