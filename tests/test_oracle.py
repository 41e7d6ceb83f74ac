"""CPU tests: the oracle against the golden vectors produced by the reference's own classes
(oracle/make_golden.py), plus self-consistency of the restated (unpinned) encodings."""
import os

import numpy as np
import pytest
import torch

from neraf_b200 import synthetic as syn
from oracle import encodings as oenc
from oracle import field as ofield
from oracle import griffinlim as ogl
from oracle import loss as oloss
from oracle import metrics as omet
from oracle import refshim


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _shape_of(meta, name):
    B, seed, C, F, T = [int(v) for v in meta]
    base = syn.RAF if name == "RAF" else syn.SOUNDSPACES
    assert (base.C, base.F, base.T) == (C, F, T)
    return base, B, seed


@pytest.mark.parametrize("name", ["RAF", "SoundSpaces"])
def test_field_forward_matches_reference(golden_dir, name):
    gold = np.load(os.path.join(golden_dir, f"field_{name}.npz"))
    shape, B, seed = _shape_of(gold["meta"], name)
    sd = syn.make_state_dict(shape, seed=seed)
    batch = syn.make_batch(shape, B, seed=seed)
    g = syn.make_grid_feature(seed)
    h = oenc.assemble_input(batch, syn.default_aabb(), shape.T, g)
    assert h.shape == (B, 1187) and h.dtype == torch.float32
    np.testing.assert_array_equal(h[:, 1024:].numpy(), gold["enc"])
    y64 = ofield.field_forward(sd, h, torch.float64)
    assert _rel(y64.numpy(), gold["y_f64"]) < 1e-13
    y32 = ofield.field_forward(sd, h, torch.float32)
    assert _rel(y32.numpy(), gold["y_f32"]) < 1e-6
    yf = ofield.field_forward_factored(sd, h[:, 1024:], g, torch.float64)
    assert _rel(yf.numpy(), gold["y_f64"]) < 1e-12
    assert y64.shape == (B, shape.C, shape.F)


@pytest.mark.parametrize("name", ["RAF", "SoundSpaces"])
def test_field_backward_and_loss_match_reference_autograd(golden_dir, name):
    gold = np.load(os.path.join(golden_dir, f"field_{name}.npz"))
    shape, B, seed = _shape_of(gold["meta"], name)
    sd = syn.make_state_dict(shape, seed=seed)
    batch = syn.make_batch(shape, B, seed=seed)
    g = syn.make_grid_feature(seed)
    enc = torch.from_numpy(gold["enc"])
    y, acts = ofield.field_forward_factored(sd, enc, g, torch.float64, keep=True)
    ld = oloss.stft_loss(y, batch["data"], "mse")
    assert abs(float(ld["audio_sc_loss"]) - float(gold["sc_f64"])) < 1e-11 * abs(float(gold["sc_f64"]))
    assert abs(float(ld["audio_mag_loss"]) - float(gold["mag_f64"])) < 1e-11 * abs(float(gold["mag_f64"]))
    dy = oloss.loss_grad(y, batch["data"], "SC+SLMSE", 1e-3)
    assert _rel(dy.numpy(), gold["dy_f64"]) < 1e-10
    grads, dgrid = ofield.field_backward(sd, acts, g, y, dy)
    assert _rel(dgrid.numpy(), gold["dgrid_f64"]) < 1e-9
    for k, v in grads.items():
        assert abs(float(v.norm()) - float(gold[f"gnorm_f64:{k}"])) < 1e-9 * float(gold[f"gnorm_f64:{k}"]), k
        if v.dim() == 2:
            assert _rel(v[:4, -8:].numpy(), gold[f"gslice_f64:{k}"]) < 1e-8, k
            assert _rel(v[-3:, :6].numpy(), gold[f"gslice0_f64:{k}"]) < 1e-8, k
        else:
            assert _rel(v[:16].numpy(), gold[f"gslice_f64:{k}"]) < 1e-8, k


def test_loss_matches_reference(golden_dir):
    gold = np.load(os.path.join(golden_dir, "loss.npz"))
    x = torch.from_numpy(gold["pred"])
    y = torch.from_numpy(gold["gt"])
    for lt, crit in (("mse", "SC+SLMSE"), ("l1", "SC+SLL1")):
        ld = oloss.stft_loss(x, y, lt)
        assert abs(float(ld["audio_sc_loss"]) - float(gold[f"sc_{lt}"])) < 1e-12 * float(gold[f"sc_{lt}"])
        assert abs(float(ld["audio_mag_loss"]) - float(gold[f"mag_{lt}"])) < 1e-12 * float(gold[f"mag_{lt}"])
        gr = oloss.loss_grad(x, y, crit, 1e-3)
        assert _rel(gr.numpy(), gold[f"grad_{lt}"]) < 1e-11
    d = oloss.loss_dict(x, y, "MSE", 1e-3)
    assert abs(float(d["audio_mse"]) - 1e-3 * float(gold["mag_mse"])) < 1e-15


def test_loss_zero_target_blows_up_like_reference():
    x = torch.zeros(2, 1, 4)
    y = torch.full((2, 1, 4), float(np.log(1e-3)))      # ym == 0 everywhere -> 0-norm denominator (no epsilon in the reference)
    ld = oloss.stft_loss(x, y, "mse")
    v = float(ld["audio_sc_loss"])
    assert (not np.isfinite(v)) or v > 1e6                # blows up instead of being regularised


@pytest.mark.parametrize("name", ["RAF", "SoundSpaces"])
def test_griffinlim_matches_torchaudio_golden(golden_dir, name):
    gold = np.load(os.path.join(golden_dir, f"griffinlim_{name}.npz"))
    n, seed, n_fft, win, hop, fs = [int(v) for v in gold["meta"]]
    mag = torch.from_numpy(gold["mag"])
    init = torch.complex(torch.from_numpy(gold["init_re"]), torch.from_numpy(gold["init_im"]))
    wave = ogl.griffinlim(mag, init, n_fft, hop, win)
    assert wave.shape == gold["wave"].shape
    assert _rel(wave.numpy(), gold["wave"]) < 2e-4          # fp32 GL, same start phase (2-5e-5)
    wave0 = ogl.griffinlim(mag, None, n_fft, hop, win)
    # the all-ones start is ill-conditioned (torch's own fp32 and fp64 runs differ by up to 4e-3): bit-equal on the host
    # that wrote the golden file, 2.9e-3 away on a host whose torch picks another FFT / ISA path
    assert _rel(wave0.numpy(), gold["wave_ones"]) < 3e-2
    w64 = ogl.griffinlim(mag.double(), init, n_fft, hop, win)
    assert _rel(w64.numpy(), gold["wave"]) < 5e-4
    for i in range(wave.shape[0]):
        for c in range(wave.shape[1]):
            assert omet.measure_edt(gold["wave"][i, c], fs=fs) == pytest.approx(gold["edt"][i, c], rel=1e-12)
            assert omet.measure_clarity(gold["wave"][i, c], fs=fs) == pytest.approx(gold["c50"][i, c], rel=1e-6)


def test_stft_istft_match_torch():
    torch.manual_seed(0)
    for n_fft, win, hop, T in ((1024, 512, 256, 60), (512, 512, 128, 20), (512, 256, 128, 9)):
        L = hop * (T - 1)
        x = torch.randn(3, L)
        w = torch.hann_window(win)
        ref = torch.stft(x, n_fft, hop, win, w, center=True, pad_mode="reflect", return_complex=True)
        mine = ogl.stft(x, n_fft, hop, win)
        assert _rel(torch.view_as_real(mine).numpy(), torch.view_as_real(ref).numpy()) < 1e-6
        back = torch.istft(ref, n_fft, hop, win, w)
        mine_b = ogl.istft(ref, n_fft, hop, win)
        assert mine_b.shape == back.shape == (3, L)
        assert _rel(mine_b.numpy(), back.numpy()) < 1e-6


def test_encoding_layout_and_rules():
    shape = syn.RAF
    batch = syn.make_batch(shape, 64, seed=3, outside_frac=0.1)
    aabb = syn.default_aabb()
    enc = oenc.encode_queries(batch, aabb, shape.T)
    assert enc.shape == (64, 163) and enc.dtype == torch.float32
    # time block: [sin 10 | cos 10 | t]
    t = batch["time_query"].float() / float(shape.T - 1)
    np.testing.assert_allclose(enc[:, 20].numpy(), t.numpy(), rtol=0, atol=0)
    f0 = enc[:, 0].double()
    np.testing.assert_allclose(f0.numpy(), np.sin(2 * np.pi * t.double().numpy()), atol=2e-6)
    # whole-vector zeroing for out-of-box microphones: raw input columns (last 3 of the mic block) are 0
    mic_n = oenc.normalize_positions(batch["mic_pose"], aabb)
    outside = ~((mic_n > 0) & (mic_n < 1)).all(-1)
    assert outside.any()
    mic_block = enc[:, 21:84]
    assert torch.all(mic_block[outside][:, 60:63] == 0)
    assert torch.all(mic_block[outside][:, :30] == 0)                 # sin(0)
    assert torch.allclose(mic_block[outside][:, 30:60], torch.ones(1))  # sin(pi/2)
    # no-grid ordering differs: [mic, src, time, rot]
    h = oenc.assemble_input(batch, aabb, shape.T, None)
    assert torch.equal(h[:, :63], enc[:, 21:84]) and torch.equal(h[:, 126:147], enc[:, :21])


def test_sh4_is_orthonormal_and_fp16():
    # Gauss-Legendre x uniform-phi quadrature of the Gram matrix (SURVEY App. D probe)
    xs, ws = np.polynomial.legendre.leggauss(32)
    phis = np.linspace(0, 2 * np.pi, 64, endpoint=False)
    dirs, wts = [], []
    for z, w in zip(xs, ws):
        r = np.sqrt(1 - z * z)
        for p in phis:
            dirs.append([r * np.cos(p), r * np.sin(p), z])
            wts.append(w * 2 * np.pi / len(phis))
    d = torch.tensor(dirs, dtype=torch.float64)
    sh = oenc.sh4_tcnn((d + 1) / 2)
    assert sh.dtype == torch.float16 and sh.shape[-1] == 16
    s = sh.double().numpy()
    gram = (s * np.array(wts)[:, None]).T @ s
    assert np.abs(gram - np.eye(16)).max() < 5e-3                      # fp16 output rounding


def test_metrics_against_closed_form():
    fs = 48000
    t60 = 0.3
    n = np.arange(int(0.6 * fs))
    h = np.exp(-6.91 * n / (t60 * fs))                                  # noiseless exponential decay: T60 exact
    assert omet.measure_rt60(h, fs, 30) == pytest.approx(t60, rel=2e-3)
    assert omet.measure_edt(h, fs) == pytest.approx(t60, rel=2e-3)
    c50 = omet.measure_clarity(h, fs=fs)
    k = int(0.05 * fs + 1)
    a = np.exp(-2 * 6.91 / (t60 * fs))
    expect = 10 * np.log10((1 - a ** k) / (a ** k - a ** len(n)))
    assert c50 == pytest.approx(expect, rel=1e-9)
    assert omet.t60_raf(h * np.random.default_rng(0).standard_normal(len(n)), fs) > 0


@pytest.mark.skipif(not refshim.available(), reason="reference tree only exists in the build container")
def test_oracle_against_live_reference_modules():
    RefField, RefLoss, helper = refshim.load()
    shape = syn.SOUNDSPACES
    sd = syn.make_state_dict(shape, seed=5)
    m = RefField(1187, 512, sound_rez=shape.C, N_frequencies=shape.F)
    m.load_state_dict(sd)
    h = torch.randn(8, 1187)
    assert _rel(ofield.field_forward(sd, h).detach().numpy(), m(h).detach().numpy()) < 1e-6
    x, y = torch.randn(8, 2, 257), torch.randn(8, 2, 257)
    ref = RefLoss(loss_type="l1")(x.double(), y.double())
    mine = oloss.stft_loss(x, y, "l1")
    assert float(ref["audio_sc_loss"]) == pytest.approx(float(mine["audio_sc_loss"]), rel=1e-12)
    assert float(ref["audio_mag_loss"]) == pytest.approx(float(mine["audio_mag_loss"]), rel=1e-12)
