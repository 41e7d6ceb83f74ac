"""Batched render / eval driver against the per-RIR calls it replaces."""
import os

import numpy as np
import pytest
import torch

from neraf_b200 import synthetic as syn
from neraf_b200.evaluator import RAFEvaluator, SoundSpacesEvaluator
from neraf_b200.model import ConstantGridFeature, NeRAFAudioModel, NeRAFAudioModelConfig
from neraf_b200.render import evaluate_rirs, render_poses

from .util import cuda, rel_fro

pytestmark = pytest.mark.gpu


def _model(shape, dev):
    cfg = NeRAFAudioModelConfig(dataset=shape.name, max_len=shape.T if shape.C == 2 else 0.32, fs=shape.fs,
                                N_freq_stft=shape.F, hop_len=shape.hop, win_len=shape.win, precision="bf16")
    model = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)))
    model.field.load_state_dict(syn.make_state_dict(shape, seed=0))
    return model.to(dev)


@pytest.mark.parametrize("shape", [syn.RAF, syn.SOUNDSPACES])
def test_render_files_equal_the_per_rir_outputs(tmp_path, shape):
    dev = cuda()
    model = _model(shape, dev)
    g = torch.Generator().manual_seed(4)
    N = 7
    mic = torch.rand(N, 3, generator=g, dtype=torch.float64) * 2 - 1
    src = torch.rand(N, 3, generator=g, dtype=torch.float64) * 2 - 1
    rot = torch.rand(N, 3, generator=g, dtype=torch.float64)
    out = render_poses(model, mic, src, rot, str(tmp_path), chunk=3, waveforms=True)
    assert out["stft"].shape == (N, shape.C, shape.F, model.max_len) and out["stft"].dtype == np.float32
    assert out["wave"].shape == (N, shape.C, shape.hop * (model.max_len - 1))
    assert sorted(os.listdir(tmp_path)) == [f"eval_{i:05d}.npy" for i in range(N)]          # NeRAF_pipeline.py:374-376
    for i in range(N):
        one = model.get_outputs_for_camera(None, None, {"mic_pose": mic[i], "source_pose": src[i], "rot": rot[i]})
        ref = one["raw_output"].permute(1, 2, 0).cpu().numpy()                              # :371
        saved = np.load(os.path.join(tmp_path, f"eval_{i:05d}.npy"))
        assert saved.shape == ref.shape and saved.dtype == np.float32
        assert np.array_equal(saved, out["stft"][i])
        assert np.max(np.abs(saved - ref)) <= 1e-5 * np.max(np.abs(ref))                    # rows do not depend on the batch


def test_evaluate_rirs_equals_per_rir_evaluation():
    dev = cuda()
    shape = syn.SOUNDSPACES
    model = _model(shape, dev)
    rng = np.random.default_rng(9)
    L_ff = shape.hop * shape.T
    items = []
    for i in range(5):
        h = (rng.standard_normal((shape.C, L_ff)) * np.exp(-6.91 * np.arange(L_ff) / (0.3 * shape.fs))).astype(np.float32)
        items.append({"data": torch.randn(shape.C, shape.F, shape.T) - 3.0, "waveform": torch.from_numpy(h),
                      "mic_pose": torch.rand(3, dtype=torch.float64), "source_pose": torch.rand(3, dtype=torch.float64),
                      "rot": torch.rand(3, dtype=torch.float64)})
    init = torch.complex(torch.rand(5, shape.C, shape.F, shape.T), torch.rand(5, shape.C, shape.F, shape.T)).to(dev)
    ev = SoundSpacesEvaluator(fs=shape.fs)
    got = evaluate_rirs(model, items, ev, chunk=2, init_phase=init)
    assert len(got) == 5
    for i, b in enumerate(items):
        y = model.query_rirs(b["mic_pose"], b["source_pose"], b["rot"])                     # (1, T, C, F)
        wav = model.istft_transform.render(y, init[i:i + 1])[0]
        ref = ev.get_full_metrics(None, None, b["waveform"], wav, wav, None, b["data"])
        assert list(got[i]) == list(ref)
        for k in ref:
            assert got[i][k] == pytest.approx(ref[k], rel=1e-6, abs=1e-9, nan_ok=True), (k, i)
    assert isinstance(RAFEvaluator(fs=48000), RAFEvaluator)
    with pytest.raises(ValueError):
        RAFEvaluator(fs=44100)
