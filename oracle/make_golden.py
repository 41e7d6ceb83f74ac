"""Generate tests/golden/*.npz from the REFERENCE's own classes (run in the build container only).

    python -m oracle.make_golden

Imports the real ``NeRAFAudioSoundField`` / ``STFTLoss`` / ``measure_edt`` /
``measure_clarity`` from /root/reference through oracle/refshim.py and
``torchaudio.transforms.GriffinLim`` (the class NeRAF_model.py:139 constructs), feeds
them the seeded synthetic inputs of neraf_b200/synthetic.py and stores inputs +
outputs.  The reference cannot travel to the GPU box; these vectors can.
Test infrastructure only.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neraf_b200 import synthetic as syn      # noqa: E402
from oracle import encodings as oenc          # noqa: E402
from oracle import refshim                    # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def field_case(shape: syn.Shape, B: int, seed: int, RefField, RefLoss):
    sd = syn.make_state_dict(shape, seed=seed)
    batch = syn.make_batch(shape, B, seed=seed)
    g = syn.make_grid_feature(seed)
    aabb = syn.default_aabb()
    h = oenc.assemble_input(batch, aabb, shape.T, g)            # (B,1187) f32 -- restated encodings (unpinned part)
    out = {}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        m = RefField(syn.N_GRID + syn.N_ENC, 512, sound_rez=shape.C, N_frequencies=shape.F)
        m.load_state_dict(sd)
        m = m.to(dt)
        gg = g.clone().to(dt).requires_grad_(True)
        enc = h[:, syn.N_GRID:].to(dt)
        hh = torch.cat([gg.expand(B, -1), enc], dim=-1)         # NeRAF_model.py:557-560
        y = m.forward(hh)                                        # reference forward
        crit = RefLoss(loss_type="mse")
        ld = crit(y, batch["data"].to(dt))                       # reference loss (NeRAF_evaluator.py:88-108)
        loss = ld["audio_sc_loss"] * 1e-1 * 1e-3 + ld["audio_mag_loss"] * 1.0 * 1e-3   # NeRAF_model.py:597-598
        y.retain_grad()
        loss.backward()
        out[f"y_{tag}"] = y.detach().numpy()
        out[f"sc_{tag}"] = ld["audio_sc_loss"].detach().numpy()
        out[f"mag_{tag}"] = ld["audio_mag_loss"].detach().numpy()
        out[f"dy_{tag}"] = y.grad.numpy()
        out[f"dgrid_{tag}"] = gg.grad.numpy()
        for name, p in m.named_parameters():
            gr = p.grad
            out[f"gnorm_{tag}:{name}"] = np.array(float(gr.norm()))
            if gr.dim() == 2:
                out[f"gslice_{tag}:{name}"] = gr[:4, -8:].numpy().copy()       # last 8 columns: the enc block of W1
                out[f"gslice0_{tag}:{name}"] = gr[-3:, :6].numpy().copy()
            else:
                out[f"gslice_{tag}:{name}"] = gr[:16].numpy().copy()
    out["enc"] = h[:, syn.N_GRID:].numpy()
    out["meta"] = np.array([B, seed, shape.C, shape.F, shape.T])
    np.savez_compressed(os.path.join(OUT, f"field_{shape.name}.npz"), **out)
    print("field", shape.name, "sc", out["sc_f64"], "mag", out["mag_f64"], "max|y|", np.abs(out["y_f64"]).max())


def loss_case(RefLoss):
    g = torch.Generator().manual_seed(7)
    x = (torch.randn(24, 2, 257, generator=g) * 2.0 - 1.0)
    y = (torch.randn(24, 2, 257, generator=g) * 2.0 - 2.0)
    out = {"pred": x.numpy(), "gt": y.numpy()}
    for lt in ("mse", "l1"):
        xx = x.double().requires_grad_(True)
        ld = RefLoss(loss_type=lt)(xx, y.double())
        (ld["audio_sc_loss"] * 1e-4 + ld["audio_mag_loss"] * 1e-3).backward()
        out[f"sc_{lt}"] = ld["audio_sc_loss"].detach().numpy()
        out[f"mag_{lt}"] = ld["audio_mag_loss"].detach().numpy()
        out[f"grad_{lt}"] = xx.grad.numpy()
        ld32 = RefLoss(loss_type=lt)(x, y)
        out[f"sc32_{lt}"] = ld32["audio_sc_loss"].numpy()
    np.savez_compressed(os.path.join(OUT, "loss.npz"), **out)
    print("loss", out["sc_mse"], out["mag_mse"], out["mag_l1"])


def griffinlim_case(shape: syn.Shape, n: int, seed: int, helper):
    import torchaudio
    rir, mag, _ = syn.make_rirs(shape, n, seed=seed)
    gl = torchaudio.transforms.GriffinLim(n_fft=shape.n_fft, win_length=shape.win, hop_length=shape.hop, power=1)
    torch.manual_seed(seed)
    init = torch.rand(mag.reshape(-1, shape.F, shape.T).size(), dtype=torch.complex64)     # what griffinlim() draws
    torch.manual_seed(seed)
    wave = gl(mag)                                                                         # rand_init=True (reference default)
    gl0 = torchaudio.transforms.GriffinLim(n_fft=shape.n_fft, win_length=shape.win, hop_length=shape.hop, power=1,
                                           rand_init=False)
    wave0 = gl0(mag)
    edt = np.array([[helper.measure_edt(w.numpy(), fs=shape.fs) for w in ws] for ws in wave])
    c50 = np.array([[helper.measure_clarity(w.numpy(), fs=shape.fs) for w in ws] for ws in wave])
    np.savez_compressed(os.path.join(OUT, f"griffinlim_{shape.name}.npz"),
                        mag=mag.numpy(), init_re=init.real.numpy().reshape(mag.shape),
                        init_im=init.imag.numpy().reshape(mag.shape), wave=wave.numpy(), wave_ones=wave0.numpy(),
                        edt=edt, c50=c50, meta=np.array([n, seed, shape.n_fft, shape.win, shape.hop, shape.fs]))
    print("griffinlim", shape.name, wave.shape, "edt", edt.ravel()[:3], "c50", c50.ravel()[:3])


def main():
    os.makedirs(OUT, exist_ok=True)
    RefField, RefLoss, helper = refshim.load()
    torch.set_num_threads(os.cpu_count())
    field_case(syn.RAF, 24, 0, RefField, RefLoss)
    field_case(syn.SOUNDSPACES, 16, 1, RefField, RefLoss)
    loss_case(RefLoss)
    griffinlim_case(syn.RAF, 3, 0, helper)
    ss = syn.Shape("SoundSpaces", 2, 257, 78, 512, 512, 128, 22050)      # office_4 length, NeRAF_config.py:43
    griffinlim_case(ss, 2, 1, helper)


if __name__ == "__main__":
    main()
