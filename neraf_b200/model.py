"""Host-side mirror of the reference audio model for the acoustic-field hot path.

``NeRAFAudioModel`` is a drop-in for the reference class of the same name (/root/reference/NeRAF/NeRAF_model.py:104-805)
as the unmodified pipeline drives it (NeRAF_pipeline.py:135-150, 186-191, 244-252, 276-281, 355-364, 438-455, 477-497):

* constructed by ``config.audio_model.setup(scene_box=..., num_train_data=..., device=...)`` -> ``Model.__init__(config,
  scene_box, num_train_data, **kwargs)`` -> ``populate_modules()`` (NeRAF_model.py:121-219);
* ``get_outputs(batch_audio)``, ``get_loss_dict(outputs, batch, metrics_dict)``, ``get_metrics_dict(outputs, batch)``,
  ``set_eval_data(...)``, ``get_outputs_for_camera(camera, obb_box, batch_audio)``,
  ``get_image_metrics_and_images(outputs, batch) -> (metrics_dict, images_dict)``, ``get_param_groups()``,
  ``update_to_step(step)``; attributes ``field``, ``grid``, ``resnet3d``, ``istft_transform``, ``evaluator``,
  ``spatial_distortion``, ``max_len``, ``mic_ch``, ``scene_box``

-- with every tensor operation of the hot path executed by the CUDA library.

Base class: nerfstudio's ``Model`` / ``ModelConfig`` when nerfstudio is importable (the class then IS a nerfstudio
model and the switch in NeRAF_config.py is one import); otherwise the restatement below of the few lines of
``nerfstudio/models/base_model.py`` the pipeline relies on [RECALLED: nerfstudio 1.x] -- same constructor, same
``populate_modules`` hook, same ``device`` property and ``device_indicator_param``.  The scene-grid feature producer is
the reference's ResNet3D-50 on the library (``config.grid_net = "resnet50"``, the default: NeRAF_model.py:185), a
learnable constant vector (``"constant"``), or any ``nn.Module`` passed as ``resnet3d=``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field as dc_field
from typing import Any, Dict, List, Optional, Tuple, Type

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .field import N_ENC, NeRAFAudioSoundField
from .griffinlim import GriffinLim
from .loss import spectral_loss

try:                                                      # the real base classes whenever nerfstudio is installed
    from nerfstudio.models.base_model import Model as _ModelBase, ModelConfig as _ModelConfigBase
    HAVE_NERFSTUDIO = True
except Exception:                                         # noqa: BLE001 - absent (this image) or broken install
    HAVE_NERFSTUDIO = False

    @dataclass
    class _ModelConfigBase:
        """nerfstudio ``InstantiateConfig``: ``setup(**kwargs)`` builds ``_target(config, **kwargs)``."""
        _target: Type = dc_field(default_factory=lambda: _ModelBase)

        def setup(self, **kwargs) -> Any:
            return self._target(self, **kwargs)

    class _ModelBase(nn.Module):
        """The part of nerfstudio's ``Model`` the pipeline uses: constructor -> ``populate_modules``; ``device`` from an
        empty parameter (it is in the reference's checkpoints as ``audio_model.device_indicator_param``)."""

        def __init__(self, config, scene_box, num_train_data: int = 0, **kwargs) -> None:
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
            pass

        def get_training_callbacks(self, training_callback_attributes=None) -> List:
            return []

        def forward(self, ray_bundle):
            return self.get_outputs(ray_bundle)

        def update_to_step(self, step: int) -> None:
            pass

        def load_model(self, loaded_state: Dict[str, Any]) -> None:
            state = {k.replace("module.", ""): v for k, v in loaded_state["model"].items()}
            self.load_state_dict(state)


@dataclass
class SceneBox:
    """Stand-in for ``nerfstudio.data.scene_box.SceneBox`` where only ``.aabb`` (2, 3) is read (NeRAF_model.py:541)."""
    aabb: torch.Tensor


@dataclass
class NeRAFAudioModelConfig(_ModelConfigBase):
    """Same fields and defaults as the reference config (NeRAF_model.py:82-101) + ``precision`` and ``grid_net``."""
    _target: Type = dc_field(default_factory=lambda: NeRAFAudioModel)
    dataset: str = "SoundSpaces"
    use_grid: bool = True
    grid_step: float = 1 / 128
    N_features: int = 1024
    use_multiple_viewing_directions: bool = True
    loss_factor: float = 1e-3
    max_len: float = 76
    W_field: int = 512
    fs: int = 22050
    criterion: str = "SC+SLMSE"
    N_freq_stft: int = 257
    hop_len: int = 128
    win_len: int = 512
    precision: str = "bf16"          # "bf16": tcgen05 tensor cores; "fp32": CUDA-core parity path
    grid_net: str = "resnet50"       # "resnet50": the reference's ResNet3D_helper (NeRAF_model.py:185); "constant": a vector


class ConstantGridFeature(nn.Module):
    """Stand-in for the ResNet3D grid-feature producer: a learnable (N_features,) vector."""

    def __init__(self, n_features: int = 1024, init: Optional[torch.Tensor] = None):
        super().__init__()
        self.feature = nn.Parameter(torch.zeros(n_features) if init is None else init.clone().float())

    def forward(self, grid: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.feature


class _QueryEncoding(nn.Module):
    """Place-holder for the reference's encoder modules (``time_encoding`` / ``position_encoding`` / ``rot_encoding``,
    NeRAF_model.py:158-171): the arithmetic lives in the library's prep kernel; what the plugin reads from these objects
    is ``get_out_dim()`` (:169-171) and ``parameters()`` (:733-736, none that hold values)."""

    def __init__(self, in_dim: int, out_dim: int, tcnn_params: bool = False):
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        if tcnn_params:
            # tcnn modules register an (empty, for a parameter-free encoding) ``params`` tensor: the key
            # ``rot_encoding.tcnn_encoding.params`` of the reference's checkpoints [RECALLED: tinycudann/modules.py]
            self.tcnn_encoding = nn.Module()
            self.tcnn_encoding.params = nn.Parameter(torch.zeros(0))

    def get_out_dim(self) -> int:
        return self.out_dim

    def forward(self, x):
        raise _lib.NerafError("the encodings are fused into the field's first layer: call NeRAFAudioModel.get_outputs or "
                              "neraf_b200.field.encode_queries")


_VIRIDIS = np.array([[0x44, 0x01, 0x54], [0x48, 0x28, 0x78], [0x3e, 0x49, 0x89], [0x31, 0x68, 0x8e], [0x26, 0x82, 0x8e],
                     [0x1f, 0x9e, 0x89], [0x35, 0xb7, 0x79], [0x6e, 0xce, 0x58], [0xb5, 0xde, 0x2b], [0xfd, 0xe7, 0x25]],
                    dtype=np.float64) / 255.0


def _viridis(v: np.ndarray) -> np.ndarray:
    """``matplotlib.cm.viridis(v)[..., :3]`` (NeRAF_model.py:781-790; display only).  matplotlib's table when it is
    importable, else a linear interpolation of the palette's ten published anchor colours."""
    try:
        from matplotlib import cm
        return cm.viridis(v)[..., :3]
    except Exception:                                     # noqa: BLE001 - matplotlib is not in the build image
        x = np.clip(np.nan_to_num(np.asarray(v, dtype=np.float64)), 0.0, 1.0) * (len(_VIRIDIS) - 1)
        i = np.minimum(x.astype(np.int64), len(_VIRIDIS) - 2)
        f = (x - i)[..., None]
        return _VIRIDIS[i] * (1 - f) + _VIRIDIS[i + 1] * f


class NeRAFAudioModel(_ModelBase):
    """NeRAF_model.py:104-805.  ``NeRAFAudioModel(config, scene_box, num_train_data, **kwargs)``; ``scene_box`` is
    anything with an ``.aabb`` (2, 3) tensor -- or that tensor itself.  Keyword extras (kept in ``self.kwargs`` like
    nerfstudio does): ``resnet3d=`` (a ready producer module), ``grid=`` (an initial (7, n, n, n) grid),
    ``process_group=`` (data parallel: global spectral-convergence sums), ``device=`` (ignored here as in nerfstudio:
    the pipeline calls ``.to(device)``)."""

    config: NeRAFAudioModelConfig

    def __init__(self, config: NeRAFAudioModelConfig, scene_box, num_train_data: int = 0, **kwargs):
        if torch.is_tensor(scene_box):
            scene_box = SceneBox(aabb=scene_box.detach().float().reshape(2, 3).clone())
        super().__init__(config, scene_box, num_train_data, **kwargs)

    def default_RAF_config(self):                           # NeRAF_model.py:109-119
        self.config.fs = 48000
        self.config.max_len = 0.32
        if self.config.fs == 48000:
            self.config.N_freq_stft, self.config.hop_len, self.config.win_len = 513, 256, 512
        elif self.config.fs == 16000:
            self.config.N_freq_stft, self.config.hop_len, self.config.win_len = 257, 128, 256

    def populate_modules(self):
        """NeRAF_model.py:121-219 (without the viewer widgets, :215-219)."""
        super().populate_modules()
        from .evaluator import RAFEvaluator, SoundSpacesEvaluator
        config = self.config
        kwargs = getattr(self, "kwargs", {})
        self.dataset = config.dataset
        if self.dataset == "RAF":
            self.default_RAF_config()
            self.max_len = int(config.max_len * config.fs) // config.hop_len
            self.mic_ch = 1
            self.evaluator = RAFEvaluator(fs=config.fs)
        else:
            self.max_len = int(config.max_len)
            self.mic_ch = 2
            self.evaluator = SoundSpacesEvaluator(fs=config.fs)
        self.use_grid = config.use_grid
        self.loss_factor = config.loss_factor
        self.criterion_name = config.criterion
        if self.criterion_name not in _lib.CRITERIA:
            raise ValueError(f"unknown criterion {self.criterion_name}")
        self.istft_transform = GriffinLim(n_fft=(config.N_freq_stft - 1) * 2, win_length=config.win_len,
                                          hop_length=config.hop_len, power=1)
        self.spatial_distortion = None                       # injected by the pipeline (NeRAF_pipeline.py:143)
        self.time_encoding = _QueryEncoding(1, 21)
        self.position_encoding = _QueryEncoding(3, 63)
        self.rot_encoding = _QueryEncoding(3, 16, tcnn_params=True)
        self.input_ch_time = self.time_encoding.get_out_dim()
        self.input_ch_pose = self.position_encoding.get_out_dim()
        self.input_ch_rot = self.rot_encoding.get_out_dim()
        self.register_buffer("aabb", torch.as_tensor(self.scene_box.aabb).detach().float().reshape(2, 3).clone(),
                             persistent=False)
        self.process_group = kwargs.get("process_group")     # DP: global spectral-convergence sums
        n_grid = config.N_features if self.use_grid else 0
        if self.use_grid:
            self.grid_size = np.array([0, 1, 0, 1, 0, 1])
            self.grid_step = config.grid_step
            self.N_features = config.N_features
            resnet3d = kwargs.get("resnet3d")
            if resnet3d is not None:
                self.resnet3d = resnet3d
            elif config.grid_net == "constant":
                self.resnet3d = ConstantGridFeature(config.N_features)
            elif config.grid_net == "resnet50":          # NeRAF_model.py:185
                from .gridnet import ResNet3D_helper
                self.resnet3d = ResNet3D_helper(in_channels=7, backbone="resnet50", pretrained=False,
                                                grid_step=config.grid_step, N_features=config.N_features,
                                                precision=config.precision)
            else:
                raise ValueError(f"unknown grid_net {config.grid_net!r}")
            self.grid_size_after_resnet = config.N_features
            self._delta = 1e-2
            self.grid_batch_i = 0
            self.grid = kwargs.get("grid")               # plain attribute like the reference (NeRAF_pipeline.py:451-455);
                                                         # None forces the reset on first use (NeRAF_model.py:203)
        self.field = NeRAFAudioSoundField(n_grid + N_ENC, config.W_field, sound_rez=self.mic_ch,
                                          N_frequencies=config.N_freq_stft, precision=config.precision)
        self.eval_source_pose = self.eval_mic_pose = self.eval_rot = self.eval_gt = None      # :209-213
        self._grid_feature_cache = None
        self._prefetch = None          # (host tensor, device copy, ready event) of batch['data'], see get_outputs
        self._copy_stream = None

    # ---- grid feature -------------------------------------------------------------------------
    def reset_grid(self, device=None) -> None:
        """NeRAF_model.py:269-277: a zero (7, n, n, n) grid over the unit cube with the voxel-centre coordinates in channels
        4-6 (colour and density, channels 0-3, are written by the pipeline's query_grid_one_batch, which needs the vision
        field and stays with the reference)."""
        device = self.device if device is None else device
        step = self.config.grid_step
        n = int(1 / step)
        self.grid = torch.zeros((7, n, n, n), dtype=torch.float32, device=device)
        axis = torch.arange(step / 2, 1, step)
        self.grid[4:] = torch.stack(torch.meshgrid(axis, axis, axis, indexing="ij"), dim=0).to(device)

    def next_grid_batch(self, batch_size: int = 4096) -> torch.Tensor:
        """The voxel centres the reference's query_grid_one_batch visits next (NeRAF_model.py:198-203, 306-311, 402-404):
        a cursor over all n^3 centres in (x, y, z) raster order, wrapping at the end.  The caller evaluates its vision
        field there and hands the result to ``write_grid_cells``."""
        if self.grid is None:
            self.reset_grid()
        if getattr(self, "_grid_centres", None) is None or self._grid_centres.shape[0] != self.grid[0].numel():
            step = self.config.grid_step
            axis = torch.arange(step / 2, 1, step)
            self._grid_centres = torch.stack(torch.meshgrid(axis, axis, axis, indexing="ij"), dim=-1).view(-1, 3)
            self.grid_batch_i = 0
        i = self.grid_batch_i
        batch = self._grid_centres[i:i + batch_size]
        self.grid_batch_i = 0 if i + batch_size >= self._grid_centres.shape[0] else i + batch_size
        return batch

    @torch.no_grad()
    def write_grid_cells(self, batch_coordinates: torch.Tensor, rgb: torch.Tensor, density: torch.Tensor,
                         rendered_rgb: bool = False, delta: float = 1e-2) -> None:
        """NeRAF_model.py:372-400: colour (sigmoid of the field's raw rgb unless it was rendered) into channels 0-2 and
        alpha = clip(1 - exp(-delta density), 0, 1) into channel 3 of the cells the coordinates fall in; points outside
        the unit cube are dropped.  Invalidates the cached grid feature."""
        if self.grid is None:
            self.reset_grid()
        dev = self.grid.device
        coords, rgb, density = batch_coordinates.to(dev), rgb.to(dev), density.to(dev)
        step = self.config.grid_step
        xs, ys, zs = ((coords[:, i] / step).int() for i in range(3))
        n = self.grid.shape
        mask = (xs >= 0) & (xs < n[1]) & (ys >= 0) & (ys < n[2]) & (zs >= 0) & (zs < n[3])
        xs, ys, zs = xs[mask].long(), ys[mask].long(), zs[mask].long()
        alpha = torch.clip(1 - torch.exp(-delta * density), 0, 1)[mask]
        color = (rgb if rendered_rgb else torch.sigmoid(rgb))[mask]
        self.grid = self.grid.detach()
        for c in range(3):
            self.grid[c, xs, ys, zs] = color[:, c].float()
        self.grid[3, xs, ys, zs] = alpha.float().reshape(-1)
        self._grid_feature_cache = None

    def grid_feature(self, sync_gradient: bool = True) -> Optional[torch.Tensor]:
        """NeRAF_model.py:554-557: resnet3d(grid[None]).flatten().  ``sync_gradient=False``: the caller hands the producer a
        feature gradient that is already summed over the ranks (GraphedTrainStep under data parallelism)."""
        if not self.use_grid:
            return None
        if self.grid is None and not isinstance(self.resnet3d, ConstantGridFeature):
            self.reset_grid()                           # NeRAF_model.py:298-299
        g = self.grid.unsqueeze(0).to(self.device) if self.grid is not None else None
        feat = self.resnet3d(g)
        if isinstance(feat, (list, tuple)):
            feat = feat[-1]
        feat = feat.flatten()
        if (sync_gradient and self.process_group is not None and feat.requires_grad
                and not isinstance(self.resnet3d, ConstantGridFeature)):
            # data parallel: the producer is replicated and its backward is linear in dg, so dg (N_features floats) is
            # summed over the ranks here and the producer's own gradients come out global on every rank
            from .distributed import sum_gradient_across_ranks
            feat = sum_gradient_across_ranks(feat, self.process_group)
        return feat

    # ---- training path --------------------------------------------------------------------------
    def get_outputs(self, batch_audio: Dict[str, torch.Tensor]) -> torch.Tensor:
        """NeRAF_model.py:531-566 -> (B, C, F) log-magnitude STFT columns (fp32)."""
        order = _lib.ORDER_TIME_MIC_SRC_ROT if self.use_grid else _lib.ORDER_MIC_SRC_TIME_ROT
        self._prefetch_target(batch_audio.get("data"))
        return self.field.forward_queries(batch_audio["time_query"], batch_audio["mic_pose"],
                                          batch_audio["source_pose"], batch_audio["rot"], self.aabb, self.max_len,
                                          self.grid_feature(), order)

    def _prefetch_target(self, data) -> None:
        """The ground-truth columns (4.2 MB at B=2048) are only needed by get_loss_dict, but the pipeline hands the same
        batch dict to get_outputs first (NeRAF_pipeline.py:188,191): start their host->device copy on a copy stream now
        so that it overlaps the field forward instead of sitting between forward and loss."""
        self._prefetch = None
        if not torch.is_tensor(data) or data.is_cuda or not data.is_pinned():
            return
        dev = self.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(self._copy_stream):
            d = data.to(device=dev, dtype=torch.float32, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._prefetch = (data, d, ev)

    def get_loss_dict(self, outputs: torch.Tensor, batch: Dict[str, torch.Tensor], metrics_dict=None):
        """NeRAF_model.py:584-600 (same keys, same weights)."""
        gt = batch["data"]
        if self._prefetch is not None and self._prefetch[0] is gt:
            _, gt, ev = self._prefetch
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            gt.record_stream(cur)
            self._prefetch = None
        if self.criterion_name == "MSE":
            _, mse = spectral_loss(outputs, gt, "MSE", 0.0, self.loss_factor, self.process_group)
            return {"audio_mse": mse}
        sc, mag = spectral_loss(outputs, gt, self.criterion_name, 1e-1 * self.loss_factor, 1.0 * self.loss_factor,
                                self.process_group)
        return {"audio_sc_loss": sc, "audio_mag_loss": mag}

    def get_metrics_dict(self, outputs: torch.Tensor, batch: Dict[str, torch.Tensor]):
        """NeRAF_model.py:568-582 (called on eval batches, NeRAF_pipeline.py:249): magnitudes clip(e^x - 1e-3, 0, 1e4) of
        prediction and target -> ``evaluator.get_stft_metrics``.  Computed where the prediction lives (the reference moves
        it to the host first)."""
        with torch.no_grad():
            predicted = outputs.detach().float()
            gt = batch["data"].to(device=predicted.device, dtype=torch.float32)
            mag_prd = torch.clip(torch.exp(predicted) - 1e-3, 0.0, 10000.0)
            mag_gt = torch.clip(torch.exp(gt) - 1e-3, 0.0, 10000.0)
            return self.evaluator.get_stft_metrics(mag_prd, mag_gt)

    def set_eval_data(self, eval_source_pose, eval_mic_pose, eval_rot, eval_gt):
        """NeRAF_model.py:602-608 (NeRAF_pipeline.py:147,461)."""
        self.eval_source_pose = eval_source_pose
        self.eval_mic_pose = eval_mic_pose
        self.eval_rot = eval_rot
        self.eval_gt = eval_gt

    def get_param_groups(self) -> Dict[str, List[nn.Parameter]]:
        """NeRAF_model.py:730-737: field + (parameter-free) encoders + the grid-feature producer."""
        params = (list(self.field.parameters()) + list(self.rot_encoding.parameters())
                  + list(self.position_encoding.parameters()) + list(self.time_encoding.parameters()))
        if self.use_grid:
            params += list(self.resnet3d.parameters())
        return {"audio_fields": params}

    def data_parallel_parameters(self):
        """The parameters whose gradients are partial sums over this rank's shard and must be all-reduced
        (``distributed.GradientAllReduce``): the field's, and a ConstantGridFeature's vector.  A replicated ResNet3D
        producer is NOT in the list: ``grid_feature`` sums dg over the ranks before it enters the producer."""
        params = list(self.field.parameters())
        if self.use_grid and isinstance(self.resnet3d, ConstantGridFeature):
            params += list(self.resnet3d.parameters())
        return params

    # ---- eval / render path -----------------------------------------------------------------------
    @torch.no_grad()
    def query_rirs(self, mic_pose: torch.Tensor, source_pose: torch.Tensor, rot: torch.Tensor,
                   grid_feature: Optional[torch.Tensor] = None) -> torch.Tensor:
        """N poses x all T time bins -> (N, T, C, F) log-STFT (the body of get_outputs_for_camera, batched;
        the grid feature is computed once for all poses instead of once per RIR, NeRAF_model.py:680-683)."""
        dev = self.device
        mic = mic_pose.reshape(-1, 3).to(dev, torch.float64)
        N, T = mic.shape[0], self.max_len
        src = source_pose.reshape(-1, 3).to(dev, torch.float64).expand(N, 3)
        r = rot.reshape(-1, 3).to(dev, torch.float64).expand(N, 3)
        tq = torch.arange(T, device=dev, dtype=torch.int64).repeat(N)
        rep = lambda x: x.repeat_interleave(T, dim=0)  # noqa: E731
        g = self.grid_feature() if grid_feature is None and self.use_grid else grid_feature
        order = _lib.ORDER_TIME_MIC_SRC_ROT if self.use_grid else _lib.ORDER_MIC_SRC_TIME_ROT
        # at most ~16k queries per launch: the activations of one launch (9704 bf16 values per query) stay a few
        # hundred MB of re-used scratch instead of growing with the number of poses
        per = max(1, 16384 // T)
        ys = []
        for lo in range(0, max(N, 1), per):
            hi = min(lo + per, N)
            ys.append(self.field.forward_queries(tq[lo * T:hi * T], rep(mic[lo:hi]), rep(src[lo:hi]), rep(r[lo:hi]),
                                                 self.aabb, self.max_len, g, order))
        y = ys[0] if len(ys) == 1 else torch.cat(ys)
        return y.view(N, T, self.mic_ch, -1)

    @torch.no_grad()
    def render_rirs(self, mic_pose, source_pose, rot, init_phase: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Loudness-map workload (viz/loudness_maps.ipynb): poses -> waveforms (N, C, hop*(T-1))."""
        return self.istft_transform.render(self.query_rirs(mic_pose, source_pose, rot), init_phase)

    @torch.no_grad()
    def get_outputs_for_camera(self, camera=None, obb_box=None, batch_audio: Optional[Dict] = None):
        """Eval branch of NeRAF_model.py:610-728 (camera=None): one RIR, keys as in the reference."""
        if camera is not None:
            raise NotImplementedError("viewer cameras need nerfstudio; pass batch_audio (the ns-eval path)")
        self.set_eval_data(batch_audio["mic_pose"], batch_audio["source_pose"], batch_audio["rot"],
                           batch_audio.get("data"))                                                  # :653 (swapped there too)
        y = self.query_rirs(batch_audio["mic_pose"], batch_audio["source_pose"], batch_audio["rot"])[0]   # (T, C, F)
        stft = {}
        for ch in range(y.shape[1]):
            stft[f"stft_ch_{ch}"] = torch.flip(y[:, ch, :].transpose(0, 1).unsqueeze(-1).cpu(), [0])
        gt = self.eval_gt
        if gt is not None:
            gt = torch.as_tensor(gt)
            for ch in range(gt.shape[0]):
                stft[f"gt_ch_{ch}"] = torch.flip(gt[ch].unsqueeze(-1).cpu(), [0])
                stft[f"comparison_ch_{ch}"] = torch.cat([stft[f"stft_ch_{ch}"], stft[f"gt_ch_{ch}"]], dim=1)
            if self.use_grid and self.grid is not None:                                              # :713-720
                stft["grid"] = self.grid[0:3].mean(dim=3).permute(1, 2, 0).to(self.device)
                stft["grid_density"] = self.grid[3].mean(dim=2).unsqueeze(-1).to(self.device)
        stft["raw_output"] = y
        return stft

    @torch.no_grad()
    def get_image_metrics_and_images(self, outputs: Dict, batch: Dict) -> Tuple[Dict[str, float], Dict[str, torch.Tensor]]:
        """NeRAF_model.py:739-805 -> ``(metrics_dict, images_dict)``: Griffin-Lim on ground truth AND prediction (one
        batched launch instead of two calls), ``self.evaluator.get_full_metrics`` with the reference's seven arguments
        (the waveforms stay on the device; the evaluator measures them there), and the display images."""
        dev = self.device
        stft = outputs["raw_output"].permute(1, 2, 0)                       # (C, F, T)
        data = torch.as_tensor(batch["data"]).to(dev, torch.float32)
        mag_prd = torch.clip(torch.exp(stft) - 1e-3, 0.0, 10000.0)
        mag_gt = torch.clip(torch.exp(data) - 1e-3, 0.0, 10000.0)
        waves = self.istft_transform(torch.stack([mag_gt, mag_prd]))         # (2, C, L)
        wav_gt = torch.as_tensor(batch["waveform"])
        metrics_dict = self.evaluator.get_full_metrics(mag_prd, mag_gt, wav_gt, waves[1], waves[0], stft, data)
        images_dict: Dict[str, torch.Tensor] = {}
        ids = [k.replace("gt_ch_", "") for k in outputs if k.startswith("gt_ch_")]
        if ids:
            lo = min(float(outputs["gt_ch_" + i].min()) for i in ids)
            hi = max(float(outputs["gt_ch_" + i].max()) for i in ids)
            span = (hi - lo) if hi > lo else 1.0
            for i in ids:
                v = _viridis(((outputs["stft_ch_" + i] - lo) / span).cpu().numpy().squeeze(-1))
                g = _viridis(((outputs["gt_ch_" + i] - lo) / span).cpu().numpy().squeeze(-1))
                images_dict["comparison_ch_" + i] = torch.from_numpy(np.concatenate([v, g], axis=1))
        if self.use_grid and "grid" in outputs:
            images_dict["grid"] = outputs["grid"]
            d = outputs["grid_density"]
            rng = float(d.max() - d.min())
            d = (d - d.min()) / (rng if rng > 0 else 1.0)
            images_dict["grid_density"] = torch.from_numpy(_viridis(d.cpu().numpy().squeeze(-1)))
        return metrics_dict, images_dict


class GraphedTrainStep:
    """CUDA-graphed get_outputs -> get_loss_dict -> backward at a fixed batch size.

    At B=2048 the step is a dozen kernels of 5-170 us: launch latency and host dispatch are first-order, so the
    whole launch sequence (bf16 re-pack of the current parameters, encodings, the job-list GEMM launches, loss,
    backward) is captured once and replayed.  ``step(batch)`` copies the batch dict (host or device tensors) into
    static device buffers and replays; parameter gradients land in ``p.grad`` (static tensors owned by the graph),
    the loss dict is returned as static 0-d tensors (``total_loss``: their sum, also formed inside the graph).

    The graphs are built straight from the C-ABI calls (no autograd nodes, no tiny scalar kernels): forward + fused
    loss | backward with the loss gradient formed inside the heads' backward kernel (``neraf_loss_grad``).  Single
    process: ONE graph.  Data parallel (``model.process_group``): the spectral loss needs its four partial sums
    all-reduced between forward and backward, and a captured NCCL collective proved fragile (ranks hung at teardown),
    so the step is TWO graphs with the 32-byte all-reduce issued eagerly between them; ``allreduce_grads()`` then sums
    the single flat gradient buffer across ranks.  With a trainable grid-feature producer (a real ResNet3D) and a
    single process the autograd calls a Trainer makes are captured instead.
    """

    def __init__(self, model: NeRAFAudioModel, example_batch: Dict[str, torch.Tensor], warmup: int = 3,
                 fused_allreduce: bool = False, overlap_allreduce: bool = False,
                 grad_dtype: torch.dtype = torch.float32, functional: Optional[bool] = None,
                 fuse_loss_sums: bool = False, exchange: str = "auto"):
        self.model = model
        # Data parallel: how the gradients (and the loss's four partial sums) cross the ranks.
        #   "kernel": by the library's own kernels over symmetric memory -- the sums inside the fused loss kernel
        #             (neraf_rank_exchange), the gradients by neraf_dp_exchange_grads running BESIDE the backward's GEMM
        #             launch -- so the whole step is ONE CUDA graph with no host-issued collective (bf16 path, bf16
        #             gradients on the wire, fp32 accumulation in the switch);
        #   "nccl":   two graphs around an eager 32-byte all-reduce + one NCCL all-reduce of the flat buffer after the
        #             backward (allreduce_grads()) -- the round-1 form, kept for A/B timing and for the fp32 path;
        #   "auto":   "kernel" when it applies (bf16 field, symmetric memory available), else "nccl".
        if exchange not in ("auto", "kernel", "nccl"):
            raise ValueError("exchange must be 'auto', 'kernel' or 'nccl'")
        self.exchange = exchange
        self.kernel_exchange = False
        # True: the loss's partial sums come out of the heads' forward epilogue (neraf_field_forward_loss_sums) -- one
        # launch and a re-read of the prediction less, but the extra exp / target loads sit in the epilogue of the LAST
        # link of the forward's dependency chain: measured 4.5 us per step slower at B=2048 (tools/ab_step.py: 378.0
        # vs 373.5 us), so the default is the separate loss_sums launch behind the forward (same numbers).
        self.fuse_loss_sums = fuse_loss_sums
        self.launches_per_step = 0        # library kernels per step (counted over the last eager warm-up step)
        # Data parallel, optional: the backward is split in two graphs; everything the first one finishes (all weight
        # gradients but dW1's: 57.5 of the 60.8 MB) is all-reduced on a communication stream WHILE the second computes
        # the last dgrad and dW1 on fewer CTAs (the collective kernel needs SMs of its own: ``comm_ctas``; set
        # NCCL_MAX_CTAS accordingly).  Measured at 2 GPUs (tools/time_dp_segments.py) it LOSES to the serial exchange
        # (700 vs 612 us fp32, 639 vs 579 us bf16): capped to 16 CTAs NCCL is slower, and the two kernels contend for
        # L2 / HBM.  Off by default.
        self.overlap_allreduce = overlap_allreduce
        self.grad_dtype = grad_dtype
        self.comm_ctas = int(os.environ.get("NERAF_COMM_CTAS", "20"))
        # Data parallel, experimental: NVLS multimem.red in the weight-gradient epilogues (the all-reduce fused into the
        # GEMMs over torch symmetric memory).  Numerically identical to backward + NCCL (tools/check_dp_nvls.py) and as
        # fast at 2 GPUs, but multimem.red delivers every rank's addend to every rank (inbound traffic grows with the
        # world size) and the pushes throttle the epilogues, so it is off by default: see DESIGN.md section 8.
        self.fused_allreduce = fused_allreduce
        self.nvls = False
        self._staging = None
        self._staged_for = None
        dev = model.device
        keys = ("time_query", "mic_pose", "source_pose", "rot", "data")
        dtypes = {"time_query": torch.int64, "mic_pose": torch.float64, "source_pose": torch.float64,
                  "rot": torch.float64, "data": torch.float32}
        # the static input tensors are views of ONE byte buffer (256-byte aligned offsets): a staged batch moves in
        # with a single device copy
        shapes = {k: tuple(example_batch[k].shape) for k in keys}
        offs, cur = {}, 0
        for k in keys:
            offs[k] = cur
            cur += (example_batch[k].numel() * torch.empty((), dtype=dtypes[k]).element_size() + 255) // 256 * 256
        self._static_bytes = torch.zeros(max(cur, 256), dtype=torch.uint8, device=dev)

        def views(buf):
            out = {}
            for k in keys:
                n = example_batch[k].numel() * torch.empty((), dtype=dtypes[k]).element_size()
                out[k] = buf[offs[k]:offs[k] + n].view(dtypes[k]).view(shapes[k])
            return out
        self._views = views
        self.static = views(self._static_bytes)
        for k in keys:
            self.static[k].copy_(example_batch[k].to(device=dev, dtype=dtypes[k]))
        self.params = [p for p in model.parameters() if p.requires_grad and p.numel() > 0]   # not the empty marker parameters
        self.group = model.process_group
        if functional is None:
            functional = True
        if not functional and self.group is not None:
            raise ValueError("the data-parallel step is built from direct library calls (functional=True)")
        if not functional:
            self._capture_autograd(warmup)      # the plugin's autograd calls, captured as they are
        else:
            self._capture_functional(warmup)

    # ---- single process: autograd inside one graph ------------------------------------------------
    def _capture_autograd(self, warmup: int):
        model, dev = self.model, self.model.device
        prev = model.field.always_repack
        model.field.always_repack = True

        def run():
            out = model.get_outputs(self.static)
            ld = model.get_loss_dict(out, self.static)
            total = sum(ld.values())
            total.backward()
            self.total_loss = total.detach()
            return ld

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                for p in self.params:
                    p.grad = None
                l0 = _lib.lib().neraf_launch_count()
                run()
                self.launches_per_step = _lib.lib().neraf_launch_count() - l0
        torch.cuda.current_stream(dev).wait_stream(side)
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.losses = run()
        model.field.always_repack = prev

    # ---- graphs of direct library calls: one (single process) or two around the loss all-reduce (data parallel) ----
    def _capture_functional(self, warmup: int):
        import ctypes as C
        import torch.distributed as dist
        model, dev = self.model, self.model.device
        world = 1 if self.group is None else dist.get_world_size(self.group)
        field = model.field
        if field.precision != "bf16" and field.precision != "fp32":
            raise ValueError("unknown precision")
        lib = _lib.lib()
        weights, biases = field._param_lists()
        # A real grid-feature producer (the ResNet3D) stays an autograd module around the field's direct calls: its
        # output is copied into a static vector the field reads, and the dg the field returns drives its backward --
        # inside the same graph (single process), or eagerly around the two graphs (data parallel: dg is formed from
        # the REDUCED db1 after the gradient exchange, so it is global and the producer's gradients need no exchange).
        self._producer = model.resnet3d if (model.use_grid and not isinstance(model.resnet3d, ConstantGridFeature)) else None
        self._feat = None
        if self._producer is not None:
            self._g_static = torch.zeros(model.config.N_features, dtype=torch.float32, device=dev)
            grid_p = self._g_static
        else:
            grid_p = model.resnet3d.feature if model.use_grid else None
        n_grid = 0 if grid_p is None else grid_p.numel()
        dims = field._dims(n_grid)
        prec = _lib.PRECISIONS[field.precision]
        B = self.static["time_query"].shape[0]
        pack_b, ws_b = C.c_size_t(), C.c_size_t()
        _lib.check(lib.neraf_field_sizes(C.byref(dims), prec, B, C.byref(pack_b), C.byref(ws_b)))
        pack = torch.empty(max(pack_b.value, 16), dtype=torch.uint8, device=dev)
        ws = torch.empty(max(ws_b.value, 16), dtype=torch.uint8, device=dev)
        out = torch.empty(B, field.sound_rez, field.N_frequencies, dtype=torch.float32, device=dev)
        self.sums = torch.zeros(5, dtype=torch.float64, device=dev)
        losses = torch.zeros(2, dtype=torch.float32, device=dev)
        # Gradient buffers.  Everything that must be summed over the ranks sits in ONE flat buffer: the weight
        # gradients, the bias gradients -- and, with a grid feature, only the per-query block of dW1 as a compact
        # (n1, 164) matrix: the grid block dW1[:, :1024] = db1 (x) g and dg = W1[:, :1024]^T db1 are linear in db1 with g
        # replicated, so they are formed AFTER the exchange from the reduced db1 (neraf_field_grid_grads), which keeps
        # 20.9 MB out of the all-reduce.
        defer = grid_p is not None and self.group is not None
        n1 = weights[0].shape[0]
        ldc = (field.in_size - n_grid + 7) // 8 * 8
        red_w = list(weights[1:]) if defer else list(weights)
        sizes = [t.numel() for t in red_w] + ([n1 * ldc] if defer else []) + [b.numel() for b in biases]
        n_weight_elems = sum(sizes) - sum(b.numel() for b in biases)
        # single process: dg sits right behind the bias gradients, so the library clears both with ONE memset node
        tail_grid = grid_p is not None and not defer
        if tail_grid:
            sizes = sizes + [grid_p.numel()]
        sizes = sizes + [3]        # slack: the fused loss kernel clears the bias gradients (/ dg) in whole 16-byte words
        # Fused all-reduce (experimental): the flat buffer lives in symmetric memory and the weight-gradient GEMMs add
        # their tiles into every rank's copy through its multicast alias.  Otherwise plain memory + one NCCL all-reduce.
        self.flat_grad, mc_ptr = None, 0
        if self.fused_allreduce and field.precision == "bf16" and world > 1:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self.flat_grad = symm_mem.empty(sum(sizes), dtype=torch.float32, device=dev)
                self._symm = symm_mem.rendezvous(self.flat_grad, self.group.group_name)
                mc_ptr = int(self._symm.multicast_ptr or 0)
            except Exception:                      # symmetric memory unavailable: plain buffer + NCCL
                self.flat_grad, mc_ptr = None, 0
        if self.flat_grad is None or mc_ptr == 0:
            self.flat_grad, mc_ptr = torch.zeros(sum(sizes), dtype=torch.float32, device=dev), 0
        self.nvls = mc_ptr != 0
        self.flat_grad.zero_()
        parts = list(torch.split(self.flat_grad, sizes))
        nw = len(red_w)
        dws = [v.view_as(t) for v, t in zip(parts[:nw], red_w)]
        compact = parts[nw] if defer else None
        dbs = [v.view_as(t) for v, t in zip(parts[nw + (1 if defer else 0):], biases)]
        n_b = len(biases)
        first_b = nw + (1 if defer else 0)
        dbs = dbs[:n_b]
        dgrid = None
        if grid_p is not None:
            dgrid = parts[first_b + n_b].view_as(grid_p) if tail_grid else torch.zeros_like(grid_p)
        if defer:
            dws = [torch.zeros_like(weights[0])] + dws
        # ---- data parallel, exchange by the library's own kernels (see __init__): the bf16 weight gradients and the fp32
        # bias gradients live in ONE symmetric-memory region (every rank maps every rank's copy; NVLS multicast alias
        # when the fabric has one), a second small symmetric buffer carries the flags
        self._xchg = None
        want_kernel = (self.group is not None and defer and field.precision == "bf16" and not self.fused_allreduce
                       and not self.overlap_allreduce and self.exchange in ("auto", "kernel")
                       and all(t.shape[1] % 8 == 0 for t in weights[1:]))
        if self.exchange == "kernel" and not want_kernel:
            raise ValueError("exchange='kernel' needs a process group, the bf16 field with a grid feature and layer "
                             "widths that are multiples of 8")
        if want_kernel:
            try:
                self._xchg = self._setup_kernel_exchange(dev, world, n_weight_elems, [b.numel() for b in biases])
            except Exception as e:                 # no symmetric memory on this fabric
                if self.exchange == "kernel":
                    raise
                import warnings
                warnings.warn(f"neraf_b200: symmetric-memory gradient exchange unavailable ({e}); using NCCL")
                self._xchg = None
        self.kernel_exchange = self._xchg is not None
        if self.kernel_exchange:                   # bias gradients accumulate straight into the exchanged region
            dbs = list(torch.split(self._xchg["bias"][:sum(b.numel() for b in biases)], [b.numel() for b in biases]))
            dbs = [v.view_as(t) for v, t in zip(dbs, biases)]
        own_grid = grid_p is not None and self._producer is None       # the constant vector is a parameter of the step
        order = list(weights) + list(biases) + ([grid_p] if own_grid else [])
        views = dws + dbs + ([dgrid] if own_grid else [])
        self._dgrid = dgrid
        for t, v in zip(order, views):
            t.grad = v
        qs = _lib.Queries()
        st = self.static
        qs.batch = B
        qs.time_query, qs.mic_pose = st["time_query"].data_ptr(), st["mic_pose"].data_ptr()
        qs.source_pose, qs.rot, qs.aabb = st["source_pose"].data_ptr(), st["rot"].data_ptr(), model.aabb.data_ptr()
        qs.time_denominator = float(model.max_len - 1.0)
        qs.order = _lib.ORDER_TIME_MIC_SRC_ROT if model.use_grid else _lib.ORDER_MIC_SRC_TIME_ROT
        qs.enc, qs.enc_ld = None, 0
        crit = _lib.CRITERIA[model.criterion_name]
        w_sc = 0.0 if model.criterion_name == "MSE" else 1e-1 * model.loss_factor
        w_mag = model.loss_factor
        n_local = out.numel()
        n_total = n_local * world
        # the loss gradient is formed inside the backward's head kernel from (out, target, reduced sums)
        lg = _lib.LossGrad()
        lg.gt, lg.n_total, lg.criterion, lg.sums = st["data"].data_ptr(), n_total, crit, self.sums.data_ptr()
        lg.w_sc, lg.w_mag = w_sc, w_mag
        lg.losses = losses.data_ptr()          # ... and so are the two loss values and their sum
        self.total_loss = torch.zeros((), dtype=torch.float32, device=dev)
        lg.total = self.total_loss.data_ptr()
        # ONE loss launch: partial sums -> grid barrier (+ the ranks' exchange of the four sums) -> gradient
        # (neraf_loss_grad.fuse_sums).  Not with the two-graph NCCL form (the host all-reduces the sums in between) and
        # not when the heads' epilogue forms the sums.
        self._sync = torch.zeros(4, dtype=torch.int32, device=dev)
        self._fused_loss = ((self.group is None or self.kernel_exchange) and not self.fuse_loss_sums
                            and (self.kernel_exchange or os.environ.get("NERAF_FUSED_LOSS", "0") == "1"))
        # single process: the two launches (loss_sums_kernel + head_backward_kernel) measured 8 us per step FASTER than the
        # one-launch form at B = 2048 (profiles/r02e_ab_tuning.txt: 341 vs 350 us; each phase of the fused kernel runs on
        # half the resident threads and the barrier's serial section sits between them), so the one-launch form is used
        # where it replaces a host-issued collective (data parallel) and is opt-in otherwise (NERAF_FUSED_LOSS=1)
        self._rank_x = None
        if self._fused_loss:
            lg.fuse_sums, lg.sync = 1, self._sync.data_ptr()
            if self.kernel_exchange:
                self._rank_x = _lib.RankExchange()
                self._rank_x.world, self._rank_x.rank = world, self._xchg["rank"]
                for r, pz in enumerate(self._xchg["signal_ptrs"]):
                    self._rank_x.peers[r] = pz
                lg.exchange = C.pointer(self._rank_x)
        w_arr, b_arr = _lib.ptr_array(weights), _lib.ptr_array(biases)
        dw_arr, db_arr = _lib.ptr_array(dws), _lib.ptr_array(dbs)
        self._keep = (pack, ws, out, losses, views, qs, w_arr, b_arr, dw_arr, db_arr, dims, lg)
        self._grad_views = list(zip(order, views))

        # bf16 exchange: the weight-gradient GEMMs store bf16 straight into the buffer that is all-reduced (same element
        # layout as the fp32 flat buffer), so the only conversion passes left are the 40 KB of bias gradients before
        # and the widening of the reduced sums after the collective
        self._flat_low, dw16_arr = None, None
        if self.kernel_exchange or (self.grad_dtype == torch.bfloat16 and defer and field.precision == "bf16" and not self.nvls
                and not self.overlap_allreduce and all(t.shape[1] % 8 == 0 for t in weights[1:])):
            self._flat_low = (self._xchg["weights16"] if self.kernel_exchange
                              else torch.zeros(sum(sizes), dtype=torch.bfloat16, device=dev))
            if self.kernel_exchange:               # view with the flat buffer's layout (weights | nothing else is read)
                sizes16 = sizes[:nw + 1]
                parts16 = list(torch.split(self._flat_low[:sum(sizes16)], sizes16))
                dw16 = [parts16[nw]] + parts16[:nw]
                dw16_arr = _lib.ptr_array(dw16)
                self._n_weight_elems = n_weight_elems
        if self._flat_low is not None and not self.kernel_exchange:
            parts16 = list(torch.split(self._flat_low, sizes))
            dw16 = [parts16[nw]] + parts16[:nw]                   # entry 0: the compact dW1 block
            dw16_arr = _lib.ptr_array(dw16)
            self._n_weight_elems = n_weight_elems
        self._keep_dw16 = dw16_arr
        mc = _lib.Multicast()
        mc.local_base, mc.multicast_base, mc.bytes = self.flat_grad.data_ptr(), mc_ptr, self.flat_grad.numel() * 4
        weight_region = self.flat_grad[:n_weight_elems]
        self._bias_region = self.flat_grad[n_weight_elems:]
        zero_stream = torch.cuda.Stream(device=dev) if self.nvls else None

        self.overlap = bool(self.overlap_allreduce and not self.nvls and field.precision == "bf16" and len(weights) > 2
                            and defer)
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        opt1, opt2 = _lib.DpOptions(), _lib.DpOptions()
        for o in (opt1, opt2):
            o.mc = C.pointer(mc) if self.nvls else None
            o.dw0_compact = _lib.ptr(compact)
            o.defer_grid_grads = 1 if defer else 0
            o.loss = C.pointer(lg)
            o.dweights_bf16 = dw16_arr
            o.zero_tail_slack = 1
        if self.kernel_exchange:
            # completion counters the backward advances + the chunk table of the exchange, in the order the backward
            # finishes the gradients: heads, trunk layers L..2, the compact dW1 block, and last the bias gradients
            L = len(weights) - field.sound_rez
            self._notify = torch.zeros(_lib.NOTIFY_COUNTERS, dtype=torch.int32, device=dev)
            self._notify_meta = [(C.c_uint32 * (L + 2))() for _ in range(3)]           # offset, count, increment (host, out)
            gx = _lib.GradExchange()
            gx.world, gx.rank, gx.max_ctas = world, self._xchg["rank"], int(os.environ.get("NERAF_COMM_CTAS_KERNEL", "0"))
            offs16 = [0]
            for n_el in sizes[:nw + 1]:
                offs16.append(offs16[-1] + n_el * 2)
            # the library's counter layout (neraf_dp_options.notify): one counter per 256 rows of every gradient matrix
            rows = [t.shape[0] for t in weights[:L]] + [sum(t.shape[0] for t in weights[L:]), B]
            n_cnt = [(r + 255) // 256 for r in rows]
            cnt_off = [sum(n_cnt[:k]) for k in range(L + 2)]
            # Row groups of ~5 MB: the largest matrix (dW of trunk layer 1, 20.9 MB = two thirds of the bytes) then travels
            # in five pieces while its remaining row blocks are still being computed -- its exchange, which gated everything
            # behind it (the workers take the chunks in order), ends ~20 us earlier.  Measured per step
            # (tools/time_dp_segments.py; profiles/r02ac_time_dp8_chunk*.txt, r02ad): 2 GPUs 442.9 (one chunk per matrix) /
            # 431.4 (7 MB) / 426.1 (5 MB) / 436.4 us (3 MB: every chunk costs a flag round trip and a load-reduce round trip
            # over NVLink, ~10 us, whatever its size); 8 GPUs 442.9 / 439.6 (11 MB) / 434.5 us (7 MB).
            target_bytes = int(os.environ.get("NERAF_EXCHANGE_CHUNK_MB", "5")) << 20

            # Push form (default): multimem.st of the bf16 sums into every rank's region, then one widening launch.
            # NERAF_EXCHANGE_PULL=1: the pull form -- a rank reduces its slice in place and every rank fetches the other
            # slices (peer loads through shared memory) straight into the fp32 .grad buffers, no widening pass.  Correct
            # (tools/check_dp_equals_single.py) but measured SLOWER at 2 GPUs (499 vs 445 us per step,
            # profiles/r02s_time_dp2_pull.txt): every chunk pays two more system fences and a flag hop (~20 us per
            # chunk, 7 chunks), which the 20 us saved behind the exchange do not buy back.
            self._pull = os.environ.get("NERAF_EXCHANGE_PULL", "0") == "1"
            flat_ptr = self.flat_grad.data_ptr()

            def matrix_chunks(slot, offset, n_rows, row_bytes, dst=None, dst_ld=0, row_elems=0):
                """Row-block groups of one gradient matrix, each ~target_bytes: they travel while the rest is computed.
                dst: fp32 destination of the sums (pull form); default: the same elements of the flat fp32 buffer."""
                groups = max(1, min(n_cnt[slot], (n_rows * row_bytes + target_bytes - 1) // target_bytes))
                out, c0 = [], 0
                for g in range(groups):
                    c1 = n_cnt[slot] * (g + 1) // groups
                    r0, r1 = c0 * 256, min(n_rows, c1 * 256)
                    off = offset + r0 * row_bytes
                    if not self._pull:
                        d = (0, 0, 0, 0)
                    elif dst is None:
                        d = (flat_ptr + off * 2, 0, 0, 0)                  # bf16 byte offset -> fp32 byte offset
                    else:
                        d = (dst + r0 * dst_ld * 4, dst_ld, row_elems, row_bytes // 2)
                    out.append((off, (r1 - r0) * row_bytes, cnt_off[slot] + c0, c1 - c0, 0) + d)
                    c0 = c1
                return out
            chunks = matrix_chunks(L, offs16[L - 1], rows[L], weights[L].shape[1] * 2)   # heads (contiguous)
            for i in range(L - 1, 0, -1):                                                # trunk layer i = red_w[i - 1]
                chunks += matrix_chunks(i, offs16[i - 1], rows[i], weights[i].shape[1] * 2)
            # the bias gradients are final when the dgrad chain ends, the compact dW1 block is the backward's last job
            chunks.append((self._xchg["bias_offset"], self._xchg["bias_bytes"], cnt_off[L + 1], n_cnt[L + 1], 1, 0, 0, 0, 0))
            # compact dW1 block: its sums go straight into columns [n_grid, n_grid + n_enc) of dW1
            # (the backward's LAST job: whatever of it travels after the last GEMM tile is exposed, so it goes in two halves)
            tail_bytes = int(float(os.environ.get("NERAF_EXCHANGE_TAIL_MB", "0.9")) * (1 << 20))
            target_bytes = min(target_bytes, max(tail_bytes, 1 << 16))
            chunks += matrix_chunks(0, offs16[nw], rows[0], ldc * 2, dst=dws[0].data_ptr() + n_grid * 4,
                                    dst_ld=weights[0].shape[1], row_elems=field.in_size - n_grid)
            if len(chunks) > _lib.MAX_EXCHANGE_CHUNKS:
                raise ValueError("too many exchange chunks: raise NERAF_EXCHANGE_CHUNK_MB")
            gx.n_chunks = len(chunks)
            gx.pull = 1 if self._pull else 0
            for c, (off, nbytes, cnt0, cnt_n, f32, dst, dst_ld, row_elems, src_ld) in enumerate(chunks):
                gx.chunks[c].offset, gx.chunks[c].bytes = off, nbytes
                gx.chunks[c].notify = self._notify.data_ptr() + 4 * cnt0
                gx.chunks[c].notify_count = cnt_n
                gx.chunks[c].f32 = f32
                gx.chunks[c].dst, gx.chunks[c].dst_ld = dst or None, dst_ld
                gx.chunks[c].row_elems, gx.chunks[c].src_ld = row_elems, src_ld
            self._exchange_chunks = chunks
            gx.multicast = self._xchg["multicast"] or None
            for r in range(world):
                gx.peers[r] = self._xchg["region_ptrs"][r]
                gx.signals[r] = self._xchg["signal_ptrs"][r]
            self._comm_state = torch.zeros(_lib.EXCHANGE_STATE_WORDS, dtype=torch.int32, device=dev)
            gx.state = self._comm_state.data_ptr()
            self._comm_trace = None
            if os.environ.get("NERAF_COMM_TRACE"):             # timeline of the exchange kernel (tools/time_dp_segments.py)
                self._comm_trace = torch.zeros(4 + 4 * len(chunks), dtype=torch.int64, device=dev)
                gx.trace = self._comm_trace.data_ptr()
            self._gx = gx
            opt1.notify = self._notify.data_ptr()
            opt1.notify_offset, opt1.notify_count, opt1.notify_increment = (
                C.cast(m, C.POINTER(C.c_uint32)) for m in self._notify_meta)
            opt1.exchange = C.pointer(gx)
        opt1.phase, opt1.max_ctas = (1 if self.overlap else 0), 0
        opt2.phase, opt2.max_ctas = 2, max(2, (sm_count - self.comm_ctas) // 2 * 2)
        # regions of the flat buffer: what the first backward graph finishes | what the second one does
        n_early = sum(t.numel() for t in (weights[1:] if defer else weights[1:]))
        self._early_region = self.flat_grad[:n_early] if defer else None
        self._late_region = self.flat_grad[n_early:] if defer else None
        if not defer:
            self.overlap = False
        self._comm_stream = torch.cuda.Stream(device=dev) if self.overlap else None

        def forward_part():
            if self.nvls:
                # every rank's copy of the weight gradients must be zero before ANY rank's backward adds into it: the
                # memset runs beside the forward, the loss's all-reduce between the two graphs is the rendezvous
                cur = torch.cuda.current_stream(dev)
                zero_stream.wait_stream(cur)
                with torch.cuda.stream(zero_stream):
                    weight_region.zero_()
            s = _lib.stream_ptr(dev)
            if self.fuse_loss_sums:
                # forward with the spectral loss's partial sums formed by the epilogue that stores the prediction
                _lib.check(lib.neraf_field_forward_loss_sums(C.byref(dims), prec, C.byref(qs), _lib.ptr(grid_p), w_arr,
                                                             b_arr, pack.data_ptr(), pack.numel(), 1, ws.data_ptr(),
                                                             ws.numel(), out.data_ptr(), 1, st["data"].data_ptr(),
                                                             self.sums.data_ptr(), s))
            else:
                # gt = NULL: the forward only clears the sums (no memset node), the loss kernel then accumulates
                _lib.check(lib.neraf_field_forward_loss_sums(C.byref(dims), prec, C.byref(qs), _lib.ptr(grid_p), w_arr,
                                                             b_arr, pack.data_ptr(), pack.numel(), 1, ws.data_ptr(),
                                                             ws.numel(), out.data_ptr(), 1, None,
                                                             self.sums.data_ptr(), s))
                if not self._fused_loss:
                    _lib.check(lib.neraf_spectral_loss_sums(out.data_ptr(), st["data"].data_ptr(), n_local,
                                                            self.sums.data_ptr(), 1, s))
            if self.nvls:
                torch.cuda.current_stream(dev).wait_stream(zero_stream)

        def backward_part():
            s = _lib.stream_ptr(dev)
            _lib.check(lib.neraf_field_backward_dp(C.byref(dims), prec, B, None, out.data_ptr(),
                                                   _lib.ptr(grid_p), w_arr, pack.data_ptr(), ws.data_ptr(), ws.numel(),
                                                   dw_arr, db_arr, _lib.ptr(dgrid), None, 0, C.byref(opt1), s))

        def backward_part2():
            _lib.check(lib.neraf_field_backward_dp(C.byref(dims), prec, B, None, out.data_ptr(),
                                                   _lib.ptr(grid_p), w_arr, pack.data_ptr(), ws.data_ptr(), ws.numel(),
                                                   dw_arr, db_arr, _lib.ptr(dgrid), None, 0, C.byref(opt2),
                                                   _lib.stream_ptr(dev)))

        def grid_part():          # after the exchange: the grid-block gradients from the REDUCED db1, dW1 block copied back
            if not defer:
                return
            if self.kernel_exchange and self._pull:
                # the exchange delivered every weight gradient as fp32 where it belongs: what is left is the grid block of
                # dW1 and dg, both from the reduced db1
                _lib.check(lib.neraf_field_grid_grads(C.byref(dims), grid_p.data_ptr(), weights[0].data_ptr(),
                                                      dbs[0].data_ptr(), None, 0, dws[0].data_ptr(), dgrid.data_ptr(),
                                                      self._grid_scratch.data_ptr(), None, None, 0, _lib.stream_ptr(dev)))
            elif self.kernel_exchange:
                # one launch: the bf16 sums every rank now holds -> the fp32 .grad views (all matrices but the compact dW1
                # block, which goes straight into dW1), and the grid block of dW1 / dg from the reduced db1
                n_widen = n_weight_elems - n1 * ldc
                _lib.check(lib.neraf_field_grid_grads(C.byref(dims), grid_p.data_ptr(), weights[0].data_ptr(),
                                                      dbs[0].data_ptr(), self._flat_low[n_widen:].data_ptr(), 1,
                                                      dws[0].data_ptr(), dgrid.data_ptr(), self._grid_scratch.data_ptr(),
                                                      self._flat_low.data_ptr(), self.flat_grad.data_ptr(), n_widen,
                                                      _lib.stream_ptr(dev)))
            else:
                _lib.check(lib.neraf_field_grid_grads(C.byref(dims), grid_p.data_ptr(), weights[0].data_ptr(),
                                                      dbs[0].data_ptr(), compact.data_ptr(), 0, dws[0].data_ptr(),
                                                      dgrid.data_ptr(), self._grid_scratch.data_ptr(), None, None, 0,
                                                      _lib.stream_ptr(dev)))
        self._grid_part = grid_part
        self._parts = (forward_part, backward_part, grid_part)      # tools/time_dp_parts.py times them one by one
        self._grid_scratch = torch.zeros(2 * max(n_grid, 1) + 4, dtype=torch.float32, device=dev)   # fp64 sums of dg + ticket

        if model.criterion_name == "MSE":
            self.losses = {"audio_mse": losses[1]}
        else:
            self.losses = {"audio_sc_loss": losses[0], "audio_mag_loss": losses[1]}

        def producer_forward():
            if self._producer is None:
                return None
            feat = model.grid_feature(sync_gradient=False)
            self._g_static.copy_(feat.detach().reshape(-1))
            return feat

        def producer_backward(feat):
            if feat is not None:
                feat.backward(dgrid.view_as(feat))
        self._producer_forward, self._producer_backward = producer_forward, producer_backward
        producer_params = [] if self._producer is None else [p for p in self._producer.parameters() if p.requires_grad]

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        if self.kernel_exchange:
            # ONE graph per step on every rank: forward | fused loss (the ranks trade the four sums inside it) | backward
            # with the gradient exchange kernel beside it | widening + grid-block gradients | producer backward
            def whole_step():
                feat = producer_forward()
                forward_part()
                backward_part()
                grid_part()
                producer_backward(feat)
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    for p in producer_params:
                        p.grad = None
                    l0 = lib.neraf_launch_count()
                    whole_step()
                    self.launches_per_step = lib.neraf_launch_count() - l0
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            for p in producer_params:
                p.grad = None
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                whole_step()
            return
        if self.group is None:
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    for p in producer_params:
                        p.grad = None
                    l0 = lib.neraf_launch_count()
                    feat = producer_forward()
                    forward_part()
                    backward_part()
                    producer_backward(feat)
                    self.launches_per_step = lib.neraf_launch_count() - l0
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            for p in producer_params:          # first accumulation inside the capture assigns: a replay overwrites
                p.grad = None
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                feat = producer_forward()
                forward_part()
                backward_part()
                producer_backward(feat)
            return
        with torch.cuda.stream(side):
            for _ in range(warmup):
                for p in producer_params:
                    p.grad = None
                l0 = lib.neraf_launch_count()
                feat = producer_forward()
                forward_part()
                dist.all_reduce(self.sums[:4], group=self.group)
                backward_part()
                if self.overlap:
                    backward_part2()
                self.launches_per_step = lib.neraf_launch_count() - l0
                dist.all_reduce(self._bias_region if self.nvls else self.flat_grad, group=self.group)
                grid_part()
                producer_backward(feat)
        for p in producer_params:
            p.grad = None
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph_fwd, self.graph_bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_fwd):
            forward_part()
        with torch.cuda.graph(self.graph_bwd):
            backward_part()
        if self.overlap:
            self.graph_bwd2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_bwd2):
                backward_part2()

    def _setup_kernel_exchange(self, dev, world: int, n_weight_elems: int, bias_sizes) -> dict:
        """Symmetric-memory buffers of the kernel exchange: one region (bf16 weight gradients in the flat buffer's order,
        then the fp32 bias gradients) and a NERAF_EXCHANGE_BYTES flag buffer; every rank's mapping of both."""
        import torch.distributed as dist
        n_bias = (sum(bias_sizes) + 3) // 4 * 4
        bias_offset = (n_weight_elems * 2 + 255) // 256 * 256
        nbytes = bias_offset + n_bias * 4
        rank = dist.get_rank(self.group)
        if world == 1:
            # a one-rank group (tests): nothing to map, the protocol runs against the local buffers
            region = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            signals = torch.zeros(_lib.EXCHANGE_BYTES, dtype=torch.uint8, device=dev)
            region_ptrs, signal_ptrs, multicast = [region.data_ptr()], [signals.data_ptr()], 0
            handles = None
        else:
            import torch.distributed._symmetric_memory as symm_mem
            region = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
            signals = symm_mem.empty(_lib.EXCHANGE_BYTES, dtype=torch.uint8, device=dev)
            h_region = symm_mem.rendezvous(region, self.group.group_name)
            h_signals = symm_mem.rendezvous(signals, self.group.group_name)
            region.zero_()
            signals.zero_()
            torch.cuda.synchronize(dev)
            dist.barrier(group=self.group)             # every rank's flags are zero before anyone raises one
            region_ptrs = [int(p) for p in h_region.buffer_ptrs]
            signal_ptrs = [int(p) for p in h_signals.buffer_ptrs]
            multicast = 0 if os.environ.get("NERAF_NO_MULTICAST") else int(h_region.multicast_ptr or 0)
            handles = (h_region, h_signals)
        return {"region": region, "signals": signals, "handles": handles, "rank": rank,
                "region_ptrs": region_ptrs, "signal_ptrs": signal_ptrs, "multicast": multicast,
                "weights16": region[:n_weight_elems * 2].view(torch.bfloat16),
                "bias": region[bias_offset:bias_offset + n_bias * 4].view(torch.float32),
                "bias_offset": bias_offset, "bias_bytes": n_bias * 4}

    def _reduce(self, region: torch.Tensor, dtype: torch.dtype) -> None:
        """Sum ``region`` of the flat gradient buffer over the ranks (NCCL), optionally through a bf16 copy."""
        import torch.distributed as dist
        if self._flat_low is not None:              # the backward wrote bf16 weight gradients: that IS the exchange format
            n_w = self._n_weight_elems              # biases (fp32 atomics) join them
            self._flat_low[n_w:].copy_(self.flat_grad[n_w:])
            dist.all_reduce(self._flat_low, group=self.group)
            self.flat_grad.copy_(self._flat_low)
            return
        if dtype == torch.float32:
            dist.all_reduce(region, group=self.group)
            return
        key = (region.data_ptr(), dtype)
        cache = self.__dict__.setdefault("_lowp", {})
        if key not in cache:
            cache[key] = torch.empty_like(region, dtype=dtype)
        low = cache[key]
        low.copy_(region)
        dist.all_reduce(low, group=self.group)
        region.copy_(low)

    def allreduce_grads(self, dtype: Optional[torch.dtype] = None) -> None:
        """Data parallel: finish the gradient exchange of the last step, then form the grid-block gradients from the
        reduced layer-1 bias gradient.

        Without overlap: one NCCL all-reduce of the flat buffer.  With overlap (default): the bulk was reduced on the
        communication stream while the second backward graph ran; what is left is the 3.4 MB that graph produced.
        ``dtype=torch.bfloat16`` halves the bytes on NVLink (30.4 MB instead of 60.8 MB): the buffer is rounded to bf16,
        summed, and widened back; the rounding (2^-9 relative per element) is below the bf16 path's own gradient error.
        """
        if self.group is None or self.kernel_exchange:      # kernel path: the graph already holds the global gradients
            return
        import torch.distributed as dist
        dtype = self.grad_dtype if dtype is None else dtype
        if self.nvls:
            # the weight gradients were summed by the NVSwitch inside the backward GEMMs; what is left is the 40 KB of
            # bias gradients (this collective is also where the ranks meet after their reductions were issued) and
            # the grid-block gradients, linear in the now-reduced db1
            dist.all_reduce(self._bias_region, group=self.group)
        elif self.overlap:
            torch.cuda.current_stream(self.model.device).wait_stream(self._comm_stream)
            self._reduce(self._late_region, torch.float32)
        else:
            self._reduce(self.flat_grad, dtype)
        self._grid_part()
        self._producer_backward(self._feat)        # dg is global now: so are the replicated producer's gradients
        self._feat = None

    def prefetch(self, batch: Dict[str, torch.Tensor]) -> None:
        """Start copying the NEXT step's batch (pinned host tensors) into staging buffers on a copy stream, so that the
        PCIe transfer (4.2 MB of target columns at B=2048: ~80 us) runs under the current step's kernels; the next
        ``step(batch)`` with the same dict then only copies staging -> static buffers on the device (~4 us).  What a
        prefetching data loader does for the reference's ``batch.to(device)`` (NeRAF_model.py:531-540)."""
        dev = self.model.device
        if self._staging is None:
            self._staging_bytes = torch.empty_like(self._static_bytes)
            self._staging = self._views(self._staging_bytes)
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staged_ev = torch.cuda.Event()
            self._consumed_ev = torch.cuda.Event()
            self._consumed_ev.record(torch.cuda.current_stream(dev))
        self._copy_stream.wait_event(self._consumed_ev)        # the previous staging -> static copy has read the buffers
        with torch.cuda.stream(self._copy_stream):
            for k, dst in self._staging.items():
                dst.copy_(batch[k], non_blocking=True)
            self._staged_ev.record(self._copy_stream)
        self._staged_for = batch

    def __call__(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if self._staged_for is not None and batch is self._staged_for:
            cur = torch.cuda.current_stream(self.model.device)
            cur.wait_event(self._staged_ev)
            self._static_bytes.copy_(self._staging_bytes, non_blocking=True)
            self._consumed_ev.record(cur)
            self._staged_for = None
        else:
            for k, dst in self.static.items():
                src = batch[k]
                if src is not dst:
                    dst.copy_(src, non_blocking=True)
        if self.group is None or self.kernel_exchange:
            self.graph.replay()
            for t, v in getattr(self, "_grad_views", ()):          # eager steps in between may have replaced .grad
                t.grad = v
        else:
            import torch.distributed as dist
            self._feat = self._producer_forward()          # eager autograd; its backward runs in allreduce_grads()
            self.graph_fwd.replay()
            dist.all_reduce(self.sums[:4], group=self.group)
            self.graph_bwd.replay()
            if self.overlap:
                cur = torch.cuda.current_stream(self.model.device)
                self._comm_stream.wait_stream(cur)
                with torch.cuda.stream(self._comm_stream):
                    self._reduce(self._early_region, self.grad_dtype)      # runs beside the second backward graph
                self.graph_bwd2.replay()
            for t, v in self._grad_views:          # eager steps in between may have replaced .grad
                t.grad = v
        return self.losses
