"""Acoustic metrics: the oracle against the reference's own functions / torchaudio on
CPU, the batched kernel against the oracle on the GPU."""
import os

import numpy as np
import pytest
import torch

from neraf_b200 import _lib
from neraf_b200 import synthetic as syn
from oracle import metrics as omet

from .util import cuda


def _rirs(shape, n, seed):
    """Decaying-noise impulse responses of a Griffin-Lim output's length, float32, a few awkward ones first."""
    rng = np.random.default_rng(seed)
    L = shape.hop * (shape.T - 1)
    t60 = rng.uniform(0.08, 0.9, n)
    h = rng.standard_normal((n, L)) * np.exp(-6.91 * np.arange(L)[None] / (t60[:, None] * shape.fs))
    h = (h * rng.uniform(0.05, 1.0, (n, 1))).astype(np.float32)
    h[0] = 0.0                                       # silent: EDT NaN, T60 -1
    h[1, L // 3:] = 0.0                              # zero tail (energy[:i_nz] shortens the curve)
    h[2, -1] = 10.0 * np.sqrt(np.sum(h[2].astype(np.float64) ** 2))       # loud last sample: < 5 dB of range, no fit
    h[3, 1:] = 0.0                                   # a single click at n = 0
    h[5, -1] = 0.1 * np.sqrt(np.sum(h[5].astype(np.float64) ** 2))        # ~20 dB of range: the 30 dB span is shortened
    return h


def test_highpass_restatement_against_torchaudio():
    import torchaudio
    h = _rirs(syn.RAF, 6, 0)[4]
    ref = torchaudio.functional.highpass_biquad(torch.from_numpy(h), syn.RAF.fs, 200.0).numpy()
    got = omet.highpass_biquad(h, syn.RAF.fs, 200.0)
    assert np.abs(got - ref).max() < 1e-4 * np.abs(ref).max()        # torchaudio filters in float32, the oracle in float64


def test_oracle_t60_is_the_reference_minus_one_on_failed_fits():
    h = _rirs(syn.SOUNDSPACES, 6, 1)
    for i in (0, 2, 3, 5):
        assert omet.t60_soundspaces(h[i], syn.SOUNDSPACES.fs) == -1.0
    assert 0.05 < omet.t60_soundspaces(h[4], syn.SOUNDSPACES.fs) < 1.5
    assert np.isnan(omet.measure_edt(h[0], syn.SOUNDSPACES.fs))


def test_metrics_fail_loudly_without_a_device():
    from neraf_b200.metrics import acoustic_metrics
    with pytest.raises(_lib.NerafError):
        acoustic_metrics(torch.zeros(2, 100), 48000.0)


def _oracle_all(h, fs, advanced):
    t60 = np.array([(omet.t60_raf if advanced else omet.t60_soundspaces)(x, fs) for x in h])
    edt = []
    for x in h:
        try:
            edt.append(omet.measure_edt(x, fs))
        except (ValueError, IndexError):             # the curve never drops 10 dB / is empty: numpy raises, the kernel reports NaN
            edt.append(float("nan"))
    with np.errstate(divide="ignore", invalid="ignore"):
        c50 = np.array([omet.measure_clarity(x, fs=fs) for x in h])
    return t60, np.array(edt), c50


def _check_against_oracle(got, h, shape, advanced):
    t60, edt, c50 = _oracle_all(h, shape.fs, advanced)
    one_sample = 1.0 / shape.fs
    # decay times are sample indices over fs: identical up to one sample where log10f's last bit decides a crossing
    assert np.array_equal(np.isnan(got["edt"]), np.isnan(edt))
    ok = ~np.isnan(edt)
    assert np.max(np.abs(got["edt"][ok] - edt[ok])) <= 6 * one_sample + 1e-12
    assert np.mean(got["edt"][ok] == edt[ok]) > 0.9
    assert np.array_equal(got["t60"] == -1.0, t60 == -1.0)
    ok = t60 != -1.0
    span = 6.0 if advanced else 2.0                                  # 60 / decay_db samples per index step
    assert np.max(np.abs(got["t60"][ok] - t60[ok])) <= (2 * span * one_sample if not advanced else 1e-2 * np.max(t60[ok]))
    # C50: float64 sums on the device against numpy's pairwise float32 sums
    fin = np.isfinite(c50)
    assert np.array_equal(np.isfinite(got["c50"]), fin)
    assert np.max(np.abs(got["c50"][fin] - c50[fin])) < 1e-4


@pytest.mark.parametrize("shape,advanced", [(syn.RAF, True), (syn.SOUNDSPACES, False)])
def test_metrics_core_on_host_matches_the_oracle(built, shape, advanced):
    """Drives neraf_b200/csrc/metrics_core.h (the code the GPU kernel executes per response) on the CPU."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "metrics_host_check")
    h = _rirs(shape, 24, 2)
    hz, decay = (200.0, 10.0) if advanced else (0.0, 30.0)
    out = subprocess.run([exe, str(h.shape[0]), str(h.shape[1]), repr(float(shape.fs)), str(hz), str(decay)],
                         input=np.ascontiguousarray(h).tobytes(), capture_output=True)
    assert out.returncode == 0
    m = np.frombuffer(out.stdout, dtype=np.float64).reshape(-1, 3)
    _check_against_oracle({"t60": m[:, 0], "edt": m[:, 1], "c50": m[:, 2]}, h, shape, advanced)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("shape,advanced", [(syn.RAF, True), (syn.SOUNDSPACES, False), (syn.RAF, False)])
def test_batched_metrics_match_the_oracle(shape, advanced):
    from neraf_b200.metrics import acoustic_metrics
    dev = cuda()
    h = _rirs(shape, 48, 2)
    got = {k: v.cpu().numpy() for k, v in acoustic_metrics(torch.from_numpy(h).to(dev), shape.fs, advanced).items()}
    _check_against_oracle(got, h, shape, advanced)


@pytest.mark.gpu
def test_reference_signatures_and_shapes():
    from neraf_b200 import metrics as M
    dev = cuda()
    shape = syn.SOUNDSPACES
    h = torch.from_numpy(_rirs(shape, 8, 3)).to(dev)
    gt, pred = h[4:6], h[6:8]                                        # (C, L) each
    t_gt, t_pred = M.compute_t60(gt, pred, shape.fs)
    e_gt, e_pred = M.evaluate_edt(pred, gt, shape.fs)
    c_gt, c_pred = M.evaluate_clarity(pred, gt, shape.fs)
    for a in (t_gt, t_pred, e_gt, e_pred, c_gt, c_pred):
        assert isinstance(a, np.ndarray) and a.shape == (2,)
    assert t_gt[0] == pytest.approx(omet.t60_soundspaces(h[4].cpu().numpy(), shape.fs), abs=4.0 / shape.fs)
    assert c_pred[1] == pytest.approx(omet.measure_clarity(h[7].cpu().numpy(), fs=shape.fs), abs=1e-4)
    m = M.acoustic_metrics(h.view(2, 4, -1), shape.fs)
    assert m["t60"].shape == (2, 4) and m["edt"].dtype == torch.float64
    assert M.acoustic_metrics(h[:0], shape.fs)["c50"].shape == (0,)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["RAF", "SoundSpaces"])
def test_evaluators_match_the_reference_evaluators(golden_dir, name):
    """Golden: the reference's RAFEvaluator / SoundSpacesEvaluator.get_full_metrics run per RIR
    (oracle/make_golden_evaluator.py); here per RIR through the same signature and for all RIRs in one batch."""
    from neraf_b200.evaluator import RAFEvaluator, SoundSpacesEvaluator
    cuda()
    z = np.load(os.path.join(golden_dir, f"evaluator_{name}.npz"))
    shape = syn.RAF if name == "RAF" else syn.SOUNDSPACES
    ev = (RAFEvaluator if name == "RAF" else SoundSpacesEvaluator)(fs=shape.fs)
    keys, rows = [str(k) for k in z["keys"]], z["rows"]
    batch = ev.get_full_metrics_batch(z["gt_ff"], z["prd"], z["log_gt"])
    tol = {"audio_T60": 0.15, "audio_T60_mean_error": 0.15, "audio_total_invalids_T60": 0.0,
           "audio_stft_error": 1e-3, "audio_EDT": 12.0 / shape.fs, "audio_C50": 1e-3}
    for i in range(rows.shape[0]):
        one = ev.get_full_metrics(None, None, z["gt_ff"][i], z["prd"][i], z["prd"][i], None, z["log_gt"][i])
        assert list(one) == keys == list(batch[i])
        for k, ref in zip(keys, rows[i]):
            for got in (one[k], batch[i][k]):
                if np.isnan(ref):
                    assert np.isnan(got), (k, i)
                else:
                    assert abs(got - ref) <= tol[k] + 1e-12, (k, i, got, ref)
