"""GPU parity tests of the fused Griffin-Lim kernel against torchaudio golden vectors, the oracle and the
acoustic yard-sticks (T60 / EDT / C50 within 1 %, BASELINE.json north_star)."""
import os

import numpy as np
import pytest
import torch

from neraf_b200 import _lib
from neraf_b200 import synthetic as syn
from neraf_b200.griffinlim import GriffinLim
from oracle import griffinlim as ogl
from oracle import metrics as omet
from tests.util import cuda, rel_fro

pytestmark = pytest.mark.gpu


def _gl(shape, **kw):
    return GriffinLim(n_fft=shape.n_fft, win_length=shape.win, hop_length=shape.hop, power=1, **kw)


@pytest.mark.parametrize("name", ["RAF", "SoundSpaces"])
def test_matches_torchaudio_golden(golden_dir, name):
    dev = cuda()
    gold = np.load(os.path.join(golden_dir, f"griffinlim_{name}.npz"))
    n, seed, n_fft, win, hop, fs = [int(v) for v in gold["meta"]]
    mag = torch.from_numpy(gold["mag"]).to(dev)
    init = torch.complex(torch.from_numpy(gold["init_re"]), torch.from_numpy(gold["init_im"])).to(dev)
    gl = GriffinLim(n_fft=n_fft, win_length=win, hop_length=hop, power=1)
    wave = gl(mag, init_phase=init)
    assert wave.shape == gold["wave"].shape
    assert rel_fro(wave, gold["wave"]) < 1e-3                      # same start phase: waveform parity
    gl0 = GriffinLim(n_fft=n_fft, win_length=win, hop_length=hop, power=1, rand_init=False)
    wave0 = gl0(mag)
    # the all-ones start is ill-conditioned: torch's own fp32 and fp64 runs differ by up to 4e-3 (measured)
    assert rel_fro(wave0, gold["wave_ones"]) < 3e-2
    w = wave.cpu().numpy()
    for i in range(w.shape[0]):
        for c in range(w.shape[1]):
            assert omet.measure_edt(w[i, c], fs=fs) == pytest.approx(gold["edt"][i, c], rel=1e-2)
            assert omet.measure_clarity(w[i, c], fs=fs) == pytest.approx(gold["c50"][i, c], rel=1e-2, abs=1e-2)


@pytest.mark.parametrize("shape", [syn.RAF, syn.SOUNDSPACES])
def test_acoustic_metrics_within_one_percent(shape):
    """64 synthetic decaying-noise RIRs: median |dT60|, |dEDT|, |dC50| vs the oracle (== torchaudio) under 1 %."""
    dev = cuda()
    n = 64 // shape.C
    _, mag, init = syn.make_rirs(shape, n, seed=5)
    wave = _gl(shape)(mag.to(dev), init_phase=init.to(dev)).cpu()
    ref = ogl.griffinlim(mag, init, shape.n_fft, shape.hop, shape.win)
    assert rel_fro(wave, ref) < 1e-3
    t60 = omet.t60_raf if shape.C == 1 else omet.t60_soundspaces
    d_t60, d_edt, d_c50 = [], [], []
    for a, b in zip(wave.reshape(-1, wave.shape[-1]).numpy(), ref.reshape(-1, ref.shape[-1]).numpy()):
        d_t60.append(abs(t60(a, shape.fs) - t60(b, shape.fs)) / abs(t60(b, shape.fs)))
        d_edt.append(abs(omet.measure_edt(a, shape.fs) - omet.measure_edt(b, shape.fs)) / omet.measure_edt(b, shape.fs))
        d_c50.append(abs(omet.measure_clarity(a, fs=shape.fs) - omet.measure_clarity(b, fs=shape.fs)) /
                     max(abs(omet.measure_clarity(b, fs=shape.fs)), 1e-3))
    assert np.median(d_t60) < 1e-2 and np.median(d_edt) < 1e-2 and np.median(d_c50) < 1e-2


@pytest.mark.parametrize("n_fft,win,hop,T", [(512, 256, 128, 9), (256, 256, 64, 33), (2048, 1024, 512, 12),
                                             (64, 64, 16, 40), (128, 100, 50, 17),
                                             # many frames per warp; an odd window offset (support rounded to pairs)
                                             (256, 256, 64, 110), (128, 101, 25, 120)])
def test_other_stft_geometries(n_fft, win, hop, T):
    dev = cuda()
    g = torch.Generator().manual_seed(n_fft + T)
    mag = torch.rand(3, n_fft // 2 + 1, T, generator=g) + 0.05
    init = torch.complex(torch.rand(mag.shape, generator=g), torch.rand(mag.shape, generator=g))
    gl = GriffinLim(n_fft=n_fft, win_length=win, hop_length=hop, power=1, n_iter=8)
    wave = gl(mag.to(dev), init_phase=init.to(dev))
    ref = ogl.griffinlim(mag, init, n_fft, hop, win, n_iter=8)
    assert wave.shape == ref.shape
    assert rel_fro(wave, ref) < 1e-3


def test_zero_iterations_is_plain_istft_and_power_two():
    dev = cuda()
    shape = syn.SOUNDSPACES
    _, mag, init = syn.make_rirs(shape, 2, seed=8)
    gl = GriffinLim(n_fft=shape.n_fft, win_length=shape.win, hop_length=shape.hop, power=2, n_iter=0)
    wave = gl((mag ** 2).to(dev), init_phase=init.to(dev)).cpu()
    ref = ogl.istft((mag * init).reshape(-1, shape.F, shape.T), shape.n_fft, shape.hop, shape.win).reshape(wave.shape)
    assert rel_fro(wave, ref) < 1e-5


def test_render_from_field_layout_equals_reference_permute_path():
    """(N, T, C, F) log-magnitudes -> waveforms == exp/clip (NeRAF_model.py:746-747) + permute + GriffinLim."""
    dev = cuda()
    shape = syn.SOUNDSPACES
    N = 5
    g = torch.Generator().manual_seed(3)
    log = torch.randn(N, shape.T, shape.C, shape.F, generator=g) * 1.5 - 3.0
    init = torch.complex(torch.rand(N, shape.C, shape.F, shape.T, generator=g),
                         torch.rand(N, shape.C, shape.F, shape.T, generator=g))
    gl = _gl(shape)
    wave = gl.render(log.to(dev), init_phase=init.to(dev))
    mag = ogl.log_to_mag(log).permute(0, 2, 3, 1).contiguous()
    ref = ogl.griffinlim(mag, init, shape.n_fft, shape.hop, shape.win)
    assert wave.shape == (N, shape.C, shape.hop * (shape.T - 1))
    assert rel_fro(wave, ref) < 1e-3
    via_forward = gl(mag.to(dev), init_phase=init.to(dev))
    # exp() evaluated by the kernel (expf) vs torch CPU differs in the last ulp; Griffin-Lim amplifies it
    assert rel_fro(via_forward, wave) < 2e-4


def test_many_signals_are_independent():
    """More signals than SMs (grid-stride loop): every waveform equals its single-signal run."""
    dev = cuda()
    shape = syn.RAF
    _, mag, init = syn.make_rirs(shape, 3, seed=1)
    reps = 120
    mag_b = mag.repeat(reps, 1, 1, 1).to(dev)
    init_b = init.repeat(reps, 1, 1, 1).to(dev)
    gl = _gl(shape, n_iter=4)
    wave = gl(mag_b, init_phase=init_b)
    single = gl(mag.to(dev), init_phase=init.to(dev))
    assert torch.equal(wave.reshape(reps, 3, -1), single.reshape(1, 3, -1).expand(reps, -1, -1))


def test_rejects_bad_arguments():
    dev = cuda()
    with pytest.raises(ValueError):
        GriffinLim(n_fft=512, momentum=1.0)
    gl = GriffinLim(n_fft=500, win_length=500, hop_length=125, power=1)
    with pytest.raises(_lib.NerafError):
        gl(torch.rand(1, 251, 10, device=dev))               # n_fft not a power of two
    gl = GriffinLim(n_fft=512, win_length=512, hop_length=128, power=1)
    with pytest.raises(ValueError):
        gl(torch.rand(1, 100, 10, device=dev))               # wrong number of bins
    with pytest.raises(_lib.NerafError):
        gl(torch.rand(1, 257, 10))                           # CPU tensor: no fallback


def test_no_signals_is_a_no_op():
    dev = torch.device("cuda:0")
    gl = GriffinLim(n_fft=1024, win_length=512, hop_length=256, power=1)
    out = gl(torch.zeros(0, 1, 513, 60, device=dev))
    assert out.shape == (0, 1, 256 * 59)
