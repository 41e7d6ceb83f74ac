"""The Python drop-in boundary: ``neraf_b200.model.NeRAFAudioModel`` driven exactly as the reference's unmodified
pipeline drives its audio model (/root/reference/NeRAF/NeRAF_pipeline.py:135-150 construction, :186-191 train step,
:244-252 eval batch, :276-281 / :355-364 eval image + ns-eval loop, :438-455 checkpoint load, :477-497 parameter groups
and state_dict), on top of a STUB of nerfstudio's ``Model`` / ``ModelConfig`` base classes.

nerfstudio is not installable in the build image, so the stub restates the few lines of
``nerfstudio/models/base_model.py`` and ``nerfstudio/configs/base_config.py`` that matter here [RECALLED: nerfstudio 1.x]:
``InstantiateConfig.setup(**kwargs) -> self._target(self, **kwargs)`` and ``Model.__init__(config, scene_box,
num_train_data, **kwargs)`` calling ``populate_modules()`` before it creates ``device_indicator_param``.  model.py picks
the base classes up with a plain ``from nerfstudio.models.base_model import Model, ModelConfig``: a second copy of the
module is loaded here with the stub in ``sys.modules``, so the class under test IS a subclass of the stub ``Model``.
"""
import importlib.util
import os
import sys
import types
from dataclasses import dataclass, field
from typing import Any, Dict, Type

import numpy as np
import pytest
import torch
import torch.nn as nn

from neraf_b200 import synthetic as syn

from .util import cuda, rel_fro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------ the stub base classes
def _stub_nerfstudio() -> Dict[str, types.ModuleType]:
    @dataclass
    class InstantiateConfig:
        _target: Type

        def setup(self, **kwargs) -> Any:
            return self._target(self, **kwargs)

    @dataclass
    class ModelConfig(InstantiateConfig):
        _target: Type = field(default_factory=lambda: Model)
        enable_collider: bool = True
        collider_params: Any = None
        loss_coefficients: Any = None
        eval_num_rays_per_chunk: int = 4096
        prompt: Any = None

    class Model(nn.Module):
        config: ModelConfig

        def __init__(self, config, scene_box, num_train_data, **kwargs) -> None:
            super().__init__()
            self.config = config
            self.scene_box = scene_box
            self.render_aabb = None
            self.num_train_data = num_train_data
            self.kwargs = kwargs
            self.collider = None
            self.populate_modules()
            self.callbacks = None
            self.device_indicator_param = nn.Parameter(torch.empty(0))

        @property
        def device(self):
            return self.device_indicator_param.device

        def populate_modules(self):
            self.stub_populated = True

        def update_to_step(self, step: int) -> None:
            self.stub_step = step

        def forward(self, ray_bundle):
            return self.get_outputs(ray_bundle)

    mods = {n: types.ModuleType(n) for n in ("nerfstudio", "nerfstudio.models", "nerfstudio.models.base_model")}
    mods["nerfstudio.models.base_model"].Model = Model
    mods["nerfstudio.models.base_model"].ModelConfig = ModelConfig
    mods["nerfstudio"].models = mods["nerfstudio.models"]
    mods["nerfstudio.models"].base_model = mods["nerfstudio.models.base_model"]
    return mods


@pytest.fixture(scope="module")
def ns():
    """(stub base_model module, neraf_b200.model loaded on top of it)."""
    mods = _stub_nerfstudio()
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    name = "neraf_b200._model_on_stub_nerfstudio"
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "neraf_b200", "model.py"))
        mod = importlib.util.module_from_spec(spec)
        mod.__package__ = "neraf_b200"
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    yield mods["nerfstudio.models.base_model"], mod
    sys.modules.pop(name, None)


@dataclass
class _SceneBox:                      # nerfstudio.data.scene_box.SceneBox: only .aabb is read (NeRAF_model.py:541)
    aabb: torch.Tensor


def _config(mod, shape, **over):
    kw = dict(dataset=shape.name, max_len=shape.T if shape.C == 2 else 76, fs=shape.fs, N_freq_stft=shape.F,
              hop_len=shape.hop, win_len=shape.win)
    kw.update(over)
    return mod.NeRAFAudioModelConfig(**kw)


def _eval_item(shape, T, seed):
    """``get_data_eval`` item (NeRAF_dataset.py:135-176): full (C, F, T) log-STFT + the waveform from file."""
    rir, mag, _ = syn.make_rirs(shape, 1, seed=seed)
    g = torch.Generator().manual_seed(seed)
    L_ff = shape.hop * T
    return {"data": torch.log(mag[0, :, :, :T] + 1e-3), "waveform": torch.nn.functional.pad(rir[0], (0, L_ff - rir.shape[-1])),
            "mic_pose": torch.rand(3, generator=g, dtype=torch.float64) * 2 - 1,
            "source_pose": torch.rand(3, generator=g, dtype=torch.float64) * 2 - 1,
            "rot": torch.tensor([1.0, 0.5, 0.5], dtype=torch.float64)}


# ------------------------------------------------------------------------------------------------ host side (no GPU)
def test_class_is_a_nerfstudio_model_and_builds_through_config_setup(ns):
    base, mod = ns
    assert mod.HAVE_NERFSTUDIO and issubclass(mod.NeRAFAudioModel, base.Model)
    assert issubclass(mod.NeRAFAudioModelConfig, base.ModelConfig)
    cfg = mod.NeRAFAudioModelConfig()
    # the reference's defaults (NeRAF_model.py:88-101), and the producer the reference builds (:185)
    ref_defaults = dict(dataset="SoundSpaces", use_grid=True, grid_step=1 / 128, N_features=1024,
                        use_multiple_viewing_directions=True, loss_factor=1e-3, max_len=76, W_field=512, fs=22050,
                        criterion="SC+SLMSE", N_freq_stft=257, hop_len=128, win_len=512)
    for k, v in ref_defaults.items():
        assert getattr(cfg, k) == v, k
    assert cfg.grid_net == "resnet50" and cfg.eval_num_rays_per_chunk == 4096          # inherited ModelConfig field
    # NeRAF_pipeline.py:135-139
    model = cfg.setup(scene_box=_SceneBox(syn.default_aabb()), num_train_data=123, device="cpu")
    assert isinstance(model, mod.NeRAFAudioModel) and model.stub_populated and model.num_train_data == 123
    assert model.kwargs == {"device": "cpu"} and model.scene_box.aabb.shape == (2, 3)
    from neraf_b200.gridnet import ResNet3D_helper
    from neraf_b200.evaluator import SoundSpacesEvaluator
    assert isinstance(model.resnet3d, ResNet3D_helper) and isinstance(model.evaluator, SoundSpacesEvaluator)
    assert model.max_len == 76 and model.mic_ch == 2 and model.grid is None and model.spatial_distortion is None
    assert model.field.in_size == 1024 + 21 + 63 * 2 + 16                            # :189
    assert (model.input_ch_time, model.input_ch_pose, model.input_ch_rot) == (21, 63, 16)
    assert model.device == torch.device("cpu")
    model.update_to_step(7)                                                          # NeRAF_pipeline.py:449
    assert model.stub_step == 7
    # NeRAF_pipeline.py:143,147: attribute injection + set_eval_data(source, mic, rot, data)
    model.spatial_distortion = "from the vision field"
    model.set_eval_data(1, 2, 3, 4)
    assert (model.eval_source_pose, model.eval_mic_pose, model.eval_rot, model.eval_gt) == (1, 2, 3, 4)
    # :477-490: one group, holding the field's and the producer's parameters
    groups = model.get_param_groups()
    assert list(groups) == ["audio_fields"]
    ids = {id(p) for p in groups["audio_fields"]}
    assert all(id(p) in ids for p in model.field.parameters()) and all(id(p) in ids for p in model.resnet3d.parameters())
    groups["audio_fields"].extend([nn.Parameter(torch.zeros(1))])                     # the pipeline extends the list (:487)
    torch.optim.Adam(groups["audio_fields"], lr=1e-4, eps=1e-15)                      # NeRAF_config.py:124
    # state_dict: the reference's names (NeRAF_field.py:41-45, NeRAF_resnet3d.py:116-176), nothing of ours added
    keys = list(model.state_dict())
    want = [f"field.soundfield.{i}.{n}" for i in range(5) for n in ("weight", "bias")]
    want += [f"field.STFT_linear.{c}.{n}" for c in range(2) for n in ("weight", "bias")]
    assert [k for k in keys if k.startswith("field.")] == want
    assert "device_indicator_param" in keys and "aabb" not in keys
    assert set(k for k in keys if k.startswith("resnet3d.")) == {"resnet3d." + k for k in syn.make_gridnet_state_dict("resnet50")}
    rest = [k for k in keys if not k.startswith(("field.", "resnet3d."))]
    assert sorted(rest) == ["device_indicator_param", "rot_encoding.tcnn_encoding.params"]


def test_raf_overrides_and_config_errors(ns):
    _, mod = ns
    cfg = mod.NeRAFAudioModelConfig(dataset="RAF", grid_net="constant")
    model = mod.NeRAFAudioModel(cfg, _SceneBox(syn.default_aabb()), 10)
    # default_RAF_config overrides whatever the config said (NeRAF_model.py:109-119,126-128)
    assert (cfg.fs, cfg.N_freq_stft, cfg.hop_len, cfg.win_len) == (48000, 513, 256, 512)
    assert model.max_len == 60 and model.mic_ch == 1 and type(model.evaluator).__name__ == "RAFEvaluator"
    with pytest.raises(ValueError):
        mod.NeRAFAudioModel(mod.NeRAFAudioModelConfig(criterion="L2", grid_net="constant"), _SceneBox(syn.default_aabb()), 0)
    # a bare aabb tensor is accepted where a scene box is expected (bench / tools)
    m2 = mod.NeRAFAudioModel(mod.NeRAFAudioModelConfig(grid_net="constant"), syn.default_aabb())
    assert torch.equal(m2.scene_box.aabb, syn.default_aabb())
    with pytest.raises(NotImplementedError):
        m2.get_outputs_for_camera(camera=object(), obb_box=None)


def test_standalone_base_matches_the_stub(ns):
    """Without nerfstudio (this image) the built-in restatement of the base class must behave like the stub."""
    from neraf_b200 import model as plain
    assert not plain.HAVE_NERFSTUDIO
    cfg = plain.NeRAFAudioModelConfig(grid_net="constant")
    m = cfg.setup(scene_box=_SceneBox(syn.default_aabb()), num_train_data=5, device="cpu")
    assert isinstance(m, plain.NeRAFAudioModel) and m.num_train_data == 5 and m.kwargs == {"device": "cpu"}
    _, mod = ns
    m_stub = mod.NeRAFAudioModelConfig(grid_net="constant").setup(scene_box=_SceneBox(syn.default_aabb()), num_train_data=5)
    assert list(m.state_dict()) == list(m_stub.state_dict())
    m.update_to_step(3)
    assert m.get_training_callbacks(None) == []


def test_viridis_fallback_is_monotone_in_luminance():
    from neraf_b200.model import _viridis
    v = _viridis(np.linspace(0, 1, 33))
    assert v.shape == (33, 3) and v.min() >= 0 and v.max() <= 1
    lum = v @ np.array([0.2126, 0.7152, 0.0722])
    assert np.all(np.diff(lum) > 0)
    assert np.allclose(_viridis(np.array([0.0, 1.0])), [[0x44 / 255, 0x01 / 255, 0x54 / 255], [0xfd / 255, 0xe7 / 255, 0x25 / 255]],
                       atol=0.02)


# ------------------------------------------------------------------------------------------------ the pipeline's calls on the GPU
@pytest.mark.gpu
@pytest.mark.parametrize("shape,grid_net", [(syn.RAF, "resnet50"), (syn.SOUNDSPACES, "constant")])
def test_pipeline_call_sequence(ns, shape, grid_net):
    from oracle import encodings as oenc, field as ofield, loss as oloss
    _, mod = ns
    dev = cuda()
    n = 64
    cfg = _config(mod, shape, grid_net=grid_net, grid_step=1 / n, precision="fp32")
    # --- NeRAF_pipeline.py:135-150
    model = cfg.setup(scene_box=_SceneBox(syn.default_aabb()), num_train_data=1000, device=dev)
    model.spatial_distortion = object()
    model.to(dev)
    T = model.max_len
    eval_items = [_eval_item(shape, T, s) for s in range(2)]
    model.set_eval_data(eval_items[0]["source_pose"], eval_items[0]["mic_pose"], eval_items[0]["rot"], eval_items[0]["data"])
    assert model.device.type == "cuda"
    sd = syn.make_state_dict(shape, seed=0)
    model.field.load_state_dict(sd)
    if grid_net == "resnet50":
        model.resnet3d.load_state_dict(syn.make_gridnet_state_dict("resnet50"))
        model.grid = syn.make_grid(n)[0]                       # what query_grid_one_batch maintains (a CPU tensor until moved)
    else:
        with torch.no_grad():
            model.resnet3d.feature.copy_(syn.make_grid_feature(0))
    model.eval()                                               # running statistics: the oracle's feature is reproducible
    # --- train step, NeRAF_pipeline.py:186-191 (+ the Trainer's backward / optimizer step)
    opt = torch.optim.Adam(model.get_param_groups()["audio_fields"], lr=1e-4, eps=1e-15)
    batch = syn.make_batch(shape, 192, seed=3)                 # host tensors, like the DataLoader's (NeRAF_datamanager.py:106-119)
    out = model.get_outputs(batch)
    assert out.shape == (192, shape.C, shape.F) and out.dtype == torch.float32 and out.requires_grad
    loss_dict = model.get_loss_dict(out, batch, {})
    assert list(loss_dict) == ["audio_sc_loss", "audio_mag_loss"] and all(v.dim() == 0 for v in loss_dict.values())
    with torch.no_grad():
        g = model.grid_feature().detach().cpu()
    enc = oenc.encode_queries(batch, syn.default_aabb(), T)
    y_ref = ofield.field_forward_factored(sd, enc, g.double(), torch.float64)
    ld_ref = oloss.loss_dict(y_ref, batch["data"])
    assert rel_fro(out.reshape(192, -1), y_ref.reshape(192, -1)) < 1e-5
    for k in loss_dict:
        assert abs(float(loss_dict[k]) - float(ld_ref[k])) < 1e-4 * abs(float(ld_ref[k])), k
    opt.zero_grad()
    sum(loss_dict.values()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.field.parameters())
    if grid_net == "resnet50":
        assert all(p.grad is not None for p in model.resnet3d.parameters())
    before = model.field.soundfield[0].weight.detach().clone()
    opt.step()
    assert not torch.equal(before, model.field.soundfield[0].weight)
    # --- eval batch, :244-252
    ebatch = syn.make_batch(shape, 64, seed=4)
    with torch.no_grad():
        eout = model.get_outputs(ebatch)
        metrics = model.get_metrics_dict(eout, ebatch)
        eloss = model.get_loss_dict(eout, ebatch, metrics)
    mag_p = torch.clip(torch.exp(eout.cpu()) - 1e-3, 0, 1e4)
    mag_g = torch.clip(torch.exp(ebatch["data"]) - 1e-3, 0, 1e4)
    assert float(metrics["audio_mag"]) == pytest.approx(float(torch.mean((mag_p - mag_g) ** 2) * 2), rel=1e-5)
    assert ("audio_spectral_loss" in metrics) == (shape.C == 1) and set(eloss) == set(loss_dict)
    # --- eval image / ns-eval loop, :276-281 and :355-387
    for item in eval_items:
        outputs = model.get_outputs_for_camera(None, None, batch_audio=item)
        metrics_dict, images_dict = model.get_image_metrics_and_images(outputs, item)
        raw = outputs["raw_output"]
        assert raw.shape == (T, shape.C, shape.F)
        audio2save = raw.permute(1, 2, 0).detach().cpu().numpy()             # :371
        assert audio2save.shape == tuple(item["data"].shape) and item["data"].shape[-1] == T      # num_rays, :380
        want = {"audio_EDT", "audio_C50", "audio_total_invalids_T60"} | (
            {"audio_T60", "audio_stft_error"} if shape.C == 1 else {"audio_T60_mean_error"})
        assert set(metrics_dict) == want and all(isinstance(v, float) for v in metrics_dict.values())
        assert "num_rays_per_sec_audio" not in metrics_dict and "fps_audio" not in metrics_dict   # the loop adds them (:381-385)
        for ch in range(shape.C):
            im = images_dict[f"comparison_ch_{ch}"]
            assert im.shape == (shape.F, 2 * T, 3) and 0.0 <= float(im.min()) and float(im.max()) <= 1.0
            assert outputs[f"stft_ch_{ch}"].shape == (shape.F, T, 1)
            assert torch.equal(outputs[f"gt_ch_{ch}"][:, :, 0], torch.flip(item["data"][ch], [0]))
        if grid_net == "resnet50":
            assert images_dict["grid"].shape == (n, n, 3) and images_dict["grid_density"].shape == (n, n, 3)
        assert model.eval_gt is item["data"]                                    # :653
    # --- checkpoint round trip, :492-497 (state_dict) and :438-455 (load_pipeline)
    state = {"audio_model." + k: v.clone() for k, v in model.state_dict().items()}
    state["audio_model.grid"] = model.grid if grid_net == "resnet50" else torch.zeros(7, 4, 4, 4)
    fresh = _config(mod, shape, grid_net=grid_net, grid_step=1 / n, precision="fp32").setup(
        scene_box=_SceneBox(syn.default_aabb()), num_train_data=1000, device=dev).to(dev)
    loaded = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state.items()}
    fresh.update_to_step(400000)
    grid = loaded.pop("audio_model.grid")
    fresh.load_state_dict({k[len("audio_model."):]: v for k, v in loaded.items()})          # strict
    fresh.grid = grid.to(dev)
    fresh.eval()
    with torch.no_grad():
        again = fresh.get_outputs(ebatch)
    assert torch.equal(again, eout)
