"""Oracle: query encodings of the acoustic field (test infrastructure only).

Restates, on CPU, what ``NeRAFAudioModel.get_outputs`` does before the MLP
(/root/reference/NeRAF/NeRAF_model.py:531-562) including the arithmetic of the
three third-party pieces it calls:

* nerfstudio ``SceneBox.get_normalized_positions``  (nerfstudio >= 0.3.0,
  pyproject.toml:6; call sites NeRAF_model.py:541-542)        -- parity UNPINNED
* nerfstudio ``NeRFEncoding`` torch implementation (call sites
  NeRAF_model.py:158-163, 548-549, 551)                        -- parity UNPINNED
* tiny-cuda-nn 1.7 ``SphericalHarmonics`` degree 4 (README.md:45; call site
  NeRAF_model.py:164-167, 550)                                 -- parity UNPINNED

None of the three is importable in the build container or on the GPU box, so
the published algorithms are restated here with the
exact dtype flow of the reference: positions stay float64 until the final
``h.float()`` (NeRAF_model.py:564), time is float32 throughout, SH is computed
in float32 and rounded to float16.
"""
from __future__ import annotations

import math

import numpy as np
import torch

NUM_FREQS = 10
MIN_FREQ_EXP = 0.0
MAX_FREQ_EXP = 8.0


def nerf_freqs() -> torch.Tensor:
    """``2 ** torch.linspace(0, 8, 10)`` as float32 (nerfstudio NeRFEncoding.pytorch_fwd)."""
    return 2 ** torch.linspace(MIN_FREQ_EXP, MAX_FREQ_EXP, NUM_FREQS)


def nerf_encoding(x: torch.Tensor) -> torch.Tensor:
    """NeRFEncoding(in_dim=D, 10, 0.0, 8.0, include_input=True) -- NeRAF_model.py:158-163.

    dtype follows ``x`` (float64 positions, float32 time); ``freqs`` is float32
    and is promoted by the product.  Output layout: [sin(s) (D*10, index d*10+k),
    sin(s + pi/2) (D*10), x (D)].
    """
    scaled = 2 * torch.pi * x
    freqs = nerf_freqs()
    s = scaled[..., None] * freqs
    s = s.view(*s.shape[:-2], -1)
    enc = torch.sin(torch.cat([s, s + torch.pi / 2.0], dim=-1))
    return torch.cat([enc, x], dim=-1)


def normalize_positions(p: torch.Tensor, aabb: torch.Tensor) -> torch.Tensor:
    """SceneBox.get_normalized_positions: (p - aabb[0]) / (aabb[1] - aabb[0]).

    ``aabb`` is float32 (NeRAF_dataparser.py:160-161); the lengths are formed in
    float32 before being promoted against the float64 positions.
    """
    lengths = aabb[1] - aabb[0]
    return (p - aabb[0]) / lengths


def zero_outside(p: torch.Tensor) -> torch.Tensor:
    """Whole-vector zeroing of out-of-box poses, strict inequalities -- NeRAF_model.py:543-546."""
    selector = ((p > 0.0) & (p < 1.0)).all(dim=-1)
    return p * selector[..., None]


_SH = [np.float32(c) for c in (
    0.28209479177387814, 0.48860251190291987, 1.0925484305920792, 0.94617469575755997,
    0.31539156525251999, 0.54627421529603959, 0.59004358992664352, 2.8906114426405538,
    0.45704579946446572, 0.3731763325901154, 1.4453057213202769)]


def sh4_tcnn(rot: torch.Tensor) -> torch.Tensor:
    """tiny-cuda-nn SphericalHarmonics degree 4 on d = 2*rot - 1, float32 math, float16 output.

    The operation order below is the contract the CUDA kernel reproduces bit for
    bit (every product/sum individually rounded to float32, no FMA contraction).
    Returns float16 (B, 16) like ``SHEncoding(levels=4, implementation="tcnn")``.
    """
    f = np.float32
    r = rot.detach().to(torch.float32).numpy().astype(np.float32)
    x = r[..., 0] * f(2.0) - f(1.0)
    y = r[..., 1] * f(2.0) - f(1.0)
    z = r[..., 2] * f(2.0) - f(1.0)
    c0, c1, c2, c3, c3b, c4, c5, c6, c7, c8, c9 = _SH
    xy, xz, yz = x * y, x * z, y * z
    x2, y2, z2 = x * x, y * y, z * z
    o = np.empty(r.shape[:-1] + (16,), dtype=np.float32)
    o[..., 0] = c0
    o[..., 1] = -c1 * y
    o[..., 2] = c1 * z
    o[..., 3] = -c1 * x
    o[..., 4] = c2 * xy
    o[..., 5] = -c2 * yz
    o[..., 6] = c3 * z2 - c3b
    o[..., 7] = -c2 * xz
    o[..., 8] = c4 * x2 - c4 * y2
    o[..., 9] = c5 * y * (f(-3.0) * x2 + y2)
    o[..., 10] = c6 * xy * z
    o[..., 11] = c7 * y * (f(1.0) - f(5.0) * z2)
    o[..., 12] = c8 * z * (f(5.0) * z2 - f(3.0))
    o[..., 13] = c7 * x * (f(1.0) - f(5.0) * z2)
    o[..., 14] = c9 * z * (x2 - y2)
    o[..., 15] = c5 * x * (-x2 + f(3.0) * y2)
    return torch.from_numpy(o.astype(np.float16))


def time_feature(time_query: torch.Tensor, max_len: int) -> torch.Tensor:
    """NeRAF_model.py:533-535: float32(time_query) / float(max_len - 1), shape (B,1)."""
    return (time_query.float() / float(max_len - 1.0)).unsqueeze(-1)


def encode_queries(batch: dict, aabb: torch.Tensor, max_len: int) -> torch.Tensor:
    """The 163 per-query columns of ``h`` in the reference order [time 21, mic 63, src 63, rot 16].

    NeRAF_model.py:533-551 and the ``cat`` at :560 (float64 by type promotion),
    then ``.float()`` (:564).  Returns float32 (B, 163).
    """
    t = time_feature(batch["time_query"], max_len)
    mic = zero_outside(normalize_positions(batch["mic_pose"], aabb))
    src = zero_outside(normalize_positions(batch["source_pose"], aabb))
    enc_mic = nerf_encoding(mic)
    enc_src = nerf_encoding(src)
    enc_t = nerf_encoding(t)
    sh = sh4_tcnn(batch["rot"])
    h = torch.cat([enc_t, enc_mic, enc_src, sh], dim=-1)   # promotes to float64
    return h.float()


def assemble_input(batch: dict, aabb: torch.Tensor, max_len: int, grid_feature: torch.Tensor | None) -> torch.Tensor:
    """Full MLP input h: [grid 1024 | time | mic | src | rot] (NeRAF_model.py:557-560), float32.

    Without a grid the reference uses a DIFFERENT order [mic, src, time, rot] (:562).
    """
    enc = encode_queries(batch, aabb, max_len)
    if grid_feature is None:
        t, mic, src, sh = enc[:, :21], enc[:, 21:84], enc[:, 84:147], enc[:, 147:]
        return torch.cat([mic, src, t, sh], dim=-1)
    g = grid_feature.flatten().expand(enc.shape[0], -1)
    return torch.cat([g.to(enc.dtype), enc], dim=-1)
