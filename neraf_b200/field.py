"""Drop-in acoustic field module backed by the sm_100a kernels.

Mirrors ``NeRAFAudioSoundField`` (/root/reference/NeRAF/NeRAF_field.py:37-65): same
constructor, same parameter names and shapes (``soundfield.{0..4}``, ``STFT_linear.{c}``
fp32 ``nn.Linear`` modules) so reference checkpoints and optimizers keep working.

Two entry points:

* ``forward(h)``              -- the reference signature, any (B, in_size) input (dense path).
* ``forward_queries(...)``    -- the fused hot path used by ``NeRAFAudioModel.get_outputs``
  (NeRAF_model.py:531-566): encodings + grid-feature hoisting + MLP in one C-ABI call.

Both run entirely in the CUDA library (``precision="bf16"`` tcgen05 tensor cores,
``precision="fp32"`` CUDA-core parity path); there is no PyTorch/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib

TRUNK_WIDTHS = (5096, 2048, 1024, 1024)      # NeRAF_field.py:41-42 (hard-coded in the reference)
N_ENC = 163                                   # 21 (time) + 63 (mic) + 63 (source) + 16 (SH)


class _FieldFn(torch.autograd.Function):
    """out = field(queries | enc, grid_feature, params);  backward fills fp32 grads for every parameter."""

    @staticmethod
    def forward(ctx, module: "NeRAFAudioSoundField", q: Dict, enc: Optional[torch.Tensor],
                grid: Optional[torch.Tensor], *params: torch.Tensor):
        lib = _lib.lib()
        dev = params[0].device
        _lib.require_device(params[0], "field parameters")
        n_layers = len(params) // 2
        weights, biases = list(params[:n_layers]), list(params[n_layers:])
        n_grid = 0 if grid is None else grid.numel()
        n_enc = module.in_size - n_grid
        dims = module._dims(n_grid)
        prec = _lib.PRECISIONS[module.precision]

        qs = _lib.Queries()
        keepalive = []
        if enc is not None:
            enc_c = enc.detach()
            if enc_c.dtype != torch.float32 or enc_c.stride(-1) != 1:
                enc_c = enc_c.float().contiguous()
            B = enc_c.shape[0]
            qs.batch, qs.enc, qs.enc_ld = B, enc_c.data_ptr(), enc_c.stride(0)
            keepalive.append(enc_c)
        else:
            B = q["time_query"].shape[0]
            qs.batch = B
            qs.time_query, qs.mic_pose = q["time_query"].data_ptr(), q["mic_pose"].data_ptr()
            qs.source_pose, qs.rot, qs.aabb = q["source_pose"].data_ptr(), q["rot"].data_ptr(), q["aabb"].data_ptr()
            qs.time_denominator, qs.order = float(q["time_denominator"]), int(q["order"])
            qs.enc, qs.enc_ld = None, 0
        grid_c = None
        if grid is not None:
            grid_c = grid.detach().float().contiguous().view(-1)

        need_grad = any(ctx.needs_input_grad[2:])
        pack, repack = module._packed(dims, prec, weights, biases)
        pack_b, ws_b = C.c_size_t(), C.c_size_t()
        _lib.check(lib.neraf_field_sizes(C.byref(dims), prec, B, C.byref(pack_b), C.byref(ws_b)))
        # training keeps the workspace (saved activations) until backward; inference re-uses one grow-only buffer per
        # stream, so rendering thousands of RIRs never goes back to the allocator for its 0.3-2 GB of scratch
        ws = torch.empty(max(ws_b.value, 16), dtype=torch.uint8, device=dev) if need_grad \
            else module._inference_workspace(max(ws_b.value, 16), dev)
        out = torch.empty(B, module.sound_rez, module.N_frequencies, dtype=torch.float32, device=dev)
        w_arr, b_arr = _lib.ptr_array(weights), _lib.ptr_array(biases)
        _lib.check(lib.neraf_field_forward(C.byref(dims), prec, C.byref(qs), _lib.ptr(grid_c), w_arr, b_arr,
                                           _lib.ptr(pack), 0 if pack is None else pack.numel(), int(repack),
                                           ws.data_ptr(), ws.numel(), out.data_ptr(),
                                           1 if need_grad else 0, _lib.stream_ptr(dev)))
        if need_grad:
            ctx.module, ctx.dims, ctx.prec, ctx.B, ctx.n_enc = module, dims, prec, B, n_enc
            ctx.ws, ctx.pack, ctx.grid_c = ws, pack, grid_c
            # the bf16 operand copies are shared by every forward of this module: remember which parameter versions
            # they were derived from, the data gradient must read the same ones
            ctx.pack_sig = None if pack is None else module._pack_cache[(dims.n_grid, prec, weights[0].device)][0]
            ctx.enc_given = enc is not None
            ctx.save_for_backward(out, *params)
        return out

    @staticmethod
    def backward(ctx, dout: torch.Tensor):
        lib = _lib.lib()
        if ctx.ws is None:
            raise RuntimeError("the field's saved activations were released by a previous backward (retain_graph is not "
                               "supported: run the forward again)")
        out, *params = ctx.saved_tensors
        n_layers = len(params) // 2
        weights = list(params[:n_layers])
        dev = out.device
        if ctx.pack_sig is not None:
            hit = ctx.module._pack_cache.get((ctx.dims.n_grid, ctx.prec, weights[0].device))
            if hit is None or hit[0] != ctx.pack_sig or hit[1] is not ctx.pack:
                raise RuntimeError("the field's bf16 operand copies were re-derived from different parameter values "
                                   "between this forward and its backward (a parameter changed without its backward "
                                   "having run): run the forward again")
        dout_c = dout.contiguous().float()
        # one flat gradient buffer: weights first (their sizes keep every view 16-byte aligned), then the bias
        # gradients and dgrid back to back so that the library zeroes them with a single memset
        biases = params[n_layers:]
        want_dgrid = ctx.grid_c is not None and ctx.needs_input_grad[3]
        sizes = [w.numel() for w in weights] + [b.numel() for b in biases] + ([ctx.grid_c.numel()] if want_dgrid else [])
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
        views = torch.split(flat, sizes)
        dws = [v.view_as(w) for v, w in zip(views[:n_layers], weights)]
        dbs = [v.view_as(b) for v, b in zip(views[n_layers:2 * n_layers], biases)]
        dgrid = views[2 * n_layers] if want_dgrid else None
        denc = None
        if ctx.enc_given and ctx.needs_input_grad[2]:
            denc = torch.empty(ctx.B, ctx.n_enc, dtype=torch.float32, device=dev)
        _lib.check(lib.neraf_field_backward(
            C.byref(ctx.dims), ctx.prec, ctx.B, dout_c.data_ptr(), out.data_ptr(), _lib.ptr(ctx.grid_c),
            _lib.ptr_array(weights), _lib.ptr(ctx.pack), ctx.ws.data_ptr(), ctx.ws.numel(), _lib.ptr_array(dws),
            _lib.ptr_array(dbs), _lib.ptr(dgrid), _lib.ptr(denc), 0 if denc is None else denc.stride(0),
            _lib.stream_ptr(dev)))
        ctx.ws = None
        return (None, None, denc, dgrid, *dws, *dbs)


class NeRAFAudioSoundField(nn.Module):
    """Same constructor / parameters as the reference class; CUDA-only execution.

    Extra keyword ``precision``: ``"bf16"`` (tcgen05 tensor cores, fp32 accumulate; default) or
    ``"fp32"`` (CUDA-core parity path).
    """

    def __init__(self, in_size: int, W: int, sound_rez: int = 2, N_frequencies: int = 257, precision: str = "bf16"):
        super().__init__()
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {list(_lib.PRECISIONS)}")
        self.in_size, self.W, self.sound_rez, self.N_frequencies = in_size, W, sound_rez, N_frequencies
        self.precision = precision
        widths = [in_size, *TRUNK_WIDTHS, W]
        self.soundfield = nn.ModuleList([nn.Linear(widths[i], widths[i + 1]) for i in range(5)])
        self.STFT_linear = nn.ModuleList([nn.Linear(W, N_frequencies) for _ in range(sound_rez)])
        self._pack_cache: Dict = {}
        self._ws_cache: Dict = {}
        # Training changes the parameters every step, so the bf16 operand copies must be re-derived every
        # forward.  Eagerly this is detected through the parameters' version counters; inside a captured CUDA
        # graph (no Python on replay) set always_repack so the pack kernels are part of the graph.
        self.always_repack = False

    # ---- helpers -----------------------------------------------------------------------------
    def _param_lists(self):
        layers = list(self.soundfield) + list(self.STFT_linear)
        return [l.weight for l in layers], [l.bias for l in layers]

    def _dims(self, n_grid: int) -> _lib.FieldDims:
        return _lib.make_dims(n_grid, self.in_size - n_grid, [*TRUNK_WIDTHS, self.W], self.sound_rez,
                              self.N_frequencies)

    def _inference_workspace(self, nbytes: int, dev: torch.device) -> torch.Tensor:
        key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
        ws = self._ws_cache.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._ws_cache[key] = ws
        return ws

    def _packed(self, dims, prec, weights: List[torch.Tensor], biases: List[torch.Tensor]):
        """(buffer for the bf16 operand copies, whether neraf_field_forward must re-derive them).

        The copies are stale whenever a parameter changed (version counters; ``always_repack`` forces it for
        captured graphs); the re-pack itself runs inside neraf_field_forward, overlapped with the encodings.
        """
        if prec == _lib.PREC_FP32:
            return None, False
        key = (dims.n_grid, prec, weights[0].device)
        sig = tuple((t.data_ptr(), t._version) for t in (*weights, *biases))
        hit = self._pack_cache.get(key)
        if hit is not None and hit[0] == sig and not self.always_repack:
            return hit[1], False
        if hit is not None:
            pack = hit[1]
        else:
            pack_b, ws_b = C.c_size_t(), C.c_size_t()
            _lib.check(_lib.lib().neraf_field_sizes(C.byref(dims), prec, 0, C.byref(pack_b), C.byref(ws_b)))
            pack = torch.empty(pack_b.value, dtype=torch.uint8, device=weights[0].device)
        self._pack_cache[key] = (sig, pack)
        return pack, True

    def _apply(self, fn, *args, **kwargs):      # .to()/.cuda() invalidate the packed copies
        self._pack_cache = {}
        return super()._apply(fn, *args, **kwargs)

    # ---- reference signature ------------------------------------------------------------------
    def forward(self, h: torch.Tensor) -> torch.Tensor:
        """NeRAF_field.py:47-65: (B, in_size) -> (B, sound_rez, N_frequencies)."""
        if h.dim() != 2 or h.shape[1] != self.in_size:
            raise ValueError(f"expected (B, {self.in_size}) input, got {tuple(h.shape)}")
        _lib.require_device(h, "field input")
        ws, bs = self._param_lists()
        return _FieldFn.apply(self, None, h, None, *ws, *bs)

    # ---- fused hot path -------------------------------------------------------------------------
    def forward_queries(self, time_query: torch.Tensor, mic_pose: torch.Tensor, source_pose: torch.Tensor,
                        rot: torch.Tensor, aabb: torch.Tensor, max_len: float,
                        grid_feature: Optional[torch.Tensor] = None,
                        order: int = _lib.ORDER_TIME_MIC_SRC_ROT) -> torch.Tensor:
        """Encodings + MLP for a batch of queries (NeRAF_model.py:531-566 in one call).

        time_query int64 (B,), poses / rot float64 (B,3), aabb float32 (2,3) -- all CUDA.
        grid_feature: the flattened ResNet3D output (in_size - 163,) or None for the no-grid variant.
        """
        n_grid = 0 if grid_feature is None else grid_feature.numel()
        if self.in_size - n_grid != N_ENC:
            raise ValueError(f"in_size {self.in_size} != grid {n_grid} + {N_ENC} query-encoding columns")
        dev = self.soundfield[0].weight.device
        q = {
            "time_query": _as(time_query, torch.int64, dev), "mic_pose": _as(mic_pose, torch.float64, dev),
            "source_pose": _as(source_pose, torch.float64, dev), "rot": _as(rot, torch.float64, dev),
            "aabb": _as(aabb, torch.float32, dev), "time_denominator": float(max_len - 1.0), "order": order,
        }
        B = q["time_query"].shape[0]
        for k in ("mic_pose", "source_pose", "rot"):
            if tuple(q[k].shape) != (B, 3):
                raise ValueError(f"{k} must be ({B}, 3), got {tuple(q[k].shape)}")
        if q["aabb"].numel() != 6:
            raise ValueError("aabb must hold 6 values (2, 3)")
        ws, bs = self._param_lists()
        return _FieldFn.apply(self, q, None, grid_feature, *ws, *bs)


def _as(t: torch.Tensor, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    """Move/cast like the reference's ``.to(self.device)`` (NeRAF_model.py:533-539); contiguous."""
    return t.to(device=device, dtype=dtype, non_blocking=True).contiguous()


def encode_queries(time_query, mic_pose, source_pose, rot, aabb, max_len: float,
                   order: int = _lib.ORDER_TIME_MIC_SRC_ROT) -> torch.Tensor:
    """The 163 per-query columns of h as fp32 (B, 163), computed by the CUDA encode kernel."""
    lib = _lib.lib()
    _lib.require_device(aabb, "aabb")
    dev = aabb.device
    qs = _lib.Queries()
    tq, mic, src = _as(time_query, torch.int64, dev), _as(mic_pose, torch.float64, dev), _as(source_pose, torch.float64, dev)
    r, ab = _as(rot, torch.float64, dev), _as(aabb, torch.float32, dev)
    B = tq.shape[0]
    qs.batch, qs.time_query, qs.mic_pose, qs.source_pose = B, tq.data_ptr(), mic.data_ptr(), src.data_ptr()
    qs.rot, qs.aabb, qs.time_denominator, qs.order = r.data_ptr(), ab.data_ptr(), float(max_len - 1.0), order
    out = torch.empty(B, N_ENC, dtype=torch.float32, device=dev)
    _lib.check(lib.neraf_encode_queries(C.byref(qs), out.data_ptr(), out.stride(0), _lib.stream_ptr(dev)))
    return out
