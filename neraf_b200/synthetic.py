"""Seeded synthetic RAF- / SoundSpaces-shaped inputs (no dataset, no network).

Shapes and dtypes follow the reference batch dict
(/root/reference/NeRAF/NeRAF_dataset.py:129-130, :294-295): ``time_query`` int64
(B,), ``mic_pose`` / ``source_pose`` / ``rot`` float64 (B,3), ``data`` float32
(B,C,F).  Generation recipe: SURVEY.md section 8(d).  Pure torch-CPU, shared by
tests and bench so that the oracle and the CUDA path see identical inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch


@dataclass(frozen=True)
class Shape:
    name: str
    C: int          # microphone channels (sound_rez)
    F: int          # STFT frequency bins
    T: int          # STFT time bins (max_len)
    n_fft: int
    win: int
    hop: int
    fs: int


RAF = Shape("RAF", 1, 513, 60, 1024, 512, 256, 48000)            # NeRAF_model.py:109-119,128
SOUNDSPACES = Shape("SoundSpaces", 2, 257, 100, 512, 512, 128, 22050)  # NeRAF_model.py:97-101; T per scene NeRAF_config.py:43

N_GRID = 1024
N_ENC = 163
TRUNK = (5096, 2048, 1024, 1024)


def default_aabb() -> torch.Tensor:
    return torch.tensor([[-4.0, -1.0, -5.0], [4.0, 3.0, 5.0]], dtype=torch.float32)


def make_batch(shape: Shape, B: int, seed: int = 0, outside_frac: float = 0.01) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    aabb = default_aabb().double()
    lo, hi = aabb[0] + 1.0, aabb[1] - 1.0
    mic = lo + (hi - lo) * torch.rand(B, 3, generator=g, dtype=torch.float64)
    src = lo + (hi - lo) * torch.rand(B, 3, generator=g, dtype=torch.float64)
    n_out = int(B * outside_frac)
    if n_out:
        idx = torch.randperm(B, generator=g)[:n_out]
        mic[idx, 0] = aabb[1, 0] + 0.5          # pushed outside: whole vector must be zeroed
        idx = torch.randperm(B, generator=g)[:n_out]
        src[idx, 2] = aabb[0, 2] - 0.25
    if shape.C == 1:
        theta = torch.randint(-180, 180, (B,), generator=g).double()
    else:
        theta = 90.0 * torch.randint(0, 4, (B,), generator=g).double()
    rad = torch.deg2rad(theta)
    rot = (torch.stack([torch.cos(rad), torch.zeros_like(rad), torch.sin(rad)], dim=-1) + 1.0) / 2.0
    tq = torch.randint(0, shape.T, (B,), generator=g, dtype=torch.int64)
    decay = torch.exp(-tq.double() / shape.T * 6.9)[:, None, None]
    r = 0.1 + 2.9 * torch.rand(B, 1, 1, generator=g, dtype=torch.float64)
    noise = torch.randn(B, shape.C, shape.F, generator=g, dtype=torch.float64).abs()
    data = torch.log(noise * decay * r + 1e-3).float()
    return {"time_query": tq, "mic_pose": mic, "source_pose": src, "rot": rot, "data": data,
            "audio_idx": torch.arange(B, dtype=torch.int64)}


def make_grid_feature(seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed + 1000)
    return (torch.relu(torch.randn(N_GRID, generator=g)) * 0.5).float()


def make_state_dict(shape: Shape, W: int = 512, seed: int = 0, in_size: int = N_GRID + N_ENC) -> Dict[str, torch.Tensor]:
    """nn.Linear default init (U(+-1/sqrt(fan_in))) in the reference's state_dict naming
    (NeRAF_field.py:41-45), drawn in module-construction order from ``seed``."""
    g = torch.Generator().manual_seed(seed)
    dims = [in_size, *TRUNK, W]
    sd = {}

    def linear(prefix, fan_in, fan_out):
        bound = 1.0 / math.sqrt(fan_in)
        sd[prefix + ".weight"] = (torch.rand(fan_out, fan_in, generator=g) * 2 - 1) * bound
        sd[prefix + ".bias"] = (torch.rand(fan_out, generator=g) * 2 - 1) * bound

    for i in range(5):
        linear(f"soundfield.{i}", dims[i], dims[i + 1])
    for c in range(shape.C):
        linear(f"STFT_linear.{c}", W, shape.F)
    return sd


def make_rirs(shape: Shape, n: int, seed: int = 0):
    """Synthetic exponentially-decaying-noise RIRs, their magnitude STFTs and a shared start phase.

    h[k] = N(0,1) * exp(-6.91 k / (T60 fs)), T60 cycling over {0.15, 0.3, 0.6} s; length hop*(T-1).
    Returns (rir (n,C,L) f32, mag (n,C,F,T) f32, init_phase (n,C,F,T) complex64).
    """
    g = torch.Generator().manual_seed(seed)
    L = shape.hop * (shape.T - 1)
    t60 = torch.tensor([0.15, 0.3, 0.6], dtype=torch.float64)[torch.arange(n) % 3]
    k = torch.arange(L, dtype=torch.float64)
    env = torch.exp(-6.91 * k[None, :] / (t60[:, None] * shape.fs))
    rir = (torch.randn(n, shape.C, L, generator=g, dtype=torch.float64) * env[:, None, :]).float()
    win = torch.hann_window(shape.win, periodic=True)
    spec = torch.stft(rir.reshape(-1, L), n_fft=shape.n_fft, hop_length=shape.hop, win_length=shape.win,
                      window=win, center=True, pad_mode="reflect", return_complex=True)
    mag = spec.abs().reshape(n, shape.C, shape.F, shape.T).float()
    re = torch.rand(n, shape.C, shape.F, shape.T, generator=g)
    im = torch.rand(n, shape.C, shape.F, shape.T, generator=g)
    return rir, mag, torch.complex(re, im)
