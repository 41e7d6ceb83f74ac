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


# ---- grid-feature producer (NeRAF_resnet3d.py) ----------------------------------------------------------------------
RESNET_LAYERS = {"resnet18": ("basic", [2, 2, 2, 2]), "resnet34": ("basic", [3, 4, 6, 3]),
                 "resnet50": ("bottleneck", [3, 4, 6, 3])}


def make_grid(n: int, seed: int = 0, channels: int = 7) -> torch.Tensor:
    """A (1, 7, n, n, n) fp32 grid shaped like NeRAF_model.py:269-277 after a few query_grid_one_batch calls: colour and
    density in the first four channels of ~30 % of the voxels (zero elsewhere), voxel-centre coordinates in the last
    three."""
    g = torch.Generator().manual_seed(seed + 2000)
    grid = torch.zeros(channels, n, n, n)
    filled = (torch.rand(n, n, n, generator=g) < 0.3).float()
    grid[:channels - 3] = torch.rand(channels - 3, n, n, n, generator=g) * filled
    c = (torch.arange(n, dtype=torch.float32) + 0.5) / n
    grid[channels - 3:] = torch.stack(torch.meshgrid(c, c, c, indexing="ij"), 0)
    return grid[None]


def make_gridnet_state_dict(backbone: str = "resnet50", in_channels: int = 7, N_features: int = 1024, seed: int = 0,
                            prefix: str = "backbone_net.") -> Dict[str, torch.Tensor]:
    """A state_dict in the reference's naming (NeRAF_resnet3d.py:116-176, 265-278): xavier-normal convolution weights
    like the reference's init (:162-163); batch-norm scale / shift drawn around (1, 0) and non-trivial running
    statistics so that every term of the normalisation is exercised."""
    g = torch.Generator().manual_seed(seed + 3000)
    kind, layers = RESNET_LAYERS[backbone]
    expansion = 4 if kind == "bottleneck" else 1
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, c_in, c_out, k):
        std = math.sqrt(2.0 / ((c_in + c_out) * k ** 3))
        sd[prefix + name + ".weight"] = torch.randn(c_out, c_in, k, k, k, generator=g) * std

    def bn(name, c):
        sd[prefix + name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        sd[prefix + name + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[prefix + name + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        sd[prefix + name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[prefix + name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    conv("conv1", in_channels, 64, 5)
    bn("bn1", 64)
    in_planes = 64
    n_stages = 4 if N_features == 2048 else 3
    for s in range(n_stages):
        planes, stride = 64 * 2 ** s, (1 if s == 0 else 2)
        for b in range(layers[s]):
            p = f"layer{s + 1}.{b}."
            if kind == "bottleneck":
                conv(p + "conv1", in_planes, planes, 1); bn(p + "bn1", planes)
                conv(p + "conv2", planes, planes, 3); bn(p + "bn2", planes)
                conv(p + "conv3", planes, planes * 4, 1); bn(p + "bn3", planes * 4)
            else:
                conv(p + "conv1", in_planes, planes, 3); bn(p + "bn1", planes)
                conv(p + "conv2", planes, planes, 3); bn(p + "bn2", planes)
            if b == 0 and (stride != 1 or in_planes != planes * expansion):
                conv(p + "downsample.0", in_planes, planes * expansion, 1); bn(p + "downsample.1", planes * expansion)
            in_planes = planes * expansion
    return sd
