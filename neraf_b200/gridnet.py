"""Grid-feature producer: the reference's ResNet3D on the sm_100a library.

Mirrors ``ResNet3D_helper`` / ``ResNet3D`` / ``Bottleneck`` / ``BasicBlock`` of
/root/reference/NeRAF/NeRAF_resnet3d.py:44-299: same constructors, same parameter and buffer names and shapes
(``backbone_net.conv1.weight`` (64, C, 5, 5, 5), ``backbone_net.layer1.0.bn1.running_mean`` ...), so the reference's
checkpoints load with ``load_state_dict`` and the optimizer built by ``get_param_groups`` (NeRAF_model.py:730-737) is
unchanged.  ``forward(x)`` takes the reference's (1, C, D, H, W) grid and returns the (1, N_features, 1, 1, 1)
feature (NeRAF_model.py:554-556); the backward pass fills fp32 gradients of every parameter.

B200-native execution.  Activations live in HBM as channels-last matrices (voxels, channels), bf16 by default, so
a 1x1x1 convolution IS a GEMM on the activation matrix and a kxkxk convolution is one gather (``neraf_grid_im2col``)
plus a GEMM; with 180 GB of HBM the gathered matrices of a whole 128^3 grid (0.5 GB for the stem) are simply kept for
the weight-gradient GEMM.  Every contraction -- forward, data gradient, weight gradient (which contracts over the
voxels, read MN-major from the same row-major matrices) -- runs on the tcgen05 job-list kernel
(``neraf_gemm_bf16_jobs``); the two gradients of a convolution go out as ONE job list, with the weight gradient split
over the voxels (split-K) so that its few output tiles occupy all SMs; ``precision="fp32"`` uses the CUDA-core GEMM for
parity.  The bf16 stem runs on a channels-last copy of the grid padded from 7 to 8 channels, so every gather moves
16-byte words.  Batch normalisation is a strip reduction + one elementwise pass fused with the residual add and the
ReLU; its backward is the same two passes.

There is no PyTorch/CPU fallback: the module raises without the library or on CPU tensors.  (``GridOps`` is the one
seam: tests/test_gridnet.py substitutes a host build of the very same per-element code, csrc/gridnet_core.h, to check
this file's assembly against the reference network on the CPU.)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib

# NERAF_GRID_SCALAR=1: the plain forms everywhere (2-byte gathers, one GEMM per launch, the stem straight from the 7-channel
# grid) -- the library reads the same variable; for A/B timing and for bisecting a parity failure
PLAIN_FORMS = bool(os.environ.get("NERAF_GRID_SCALAR"))
DT_F32, DT_BF16 = 0, 1
_DT = {torch.float32: DT_F32, torch.bfloat16: DT_BF16}


class Window3d(_lib.Window3d):      # neraf_window3d
    @property
    def out_dims(self) -> Tuple[int, int, int]:
        f = lambda n: (n + 2 * self.pad - self.k) // self.stride + 1      # noqa: E731
        return f(self.in_d), f(self.in_h), f(self.in_w)


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class GridOps:
    """The C-ABI calls of the producer: ``neraf_grid_*`` plus the three GEMM forms.  All tensors are 2-D row-major."""

    prefix = "neraf_grid_"

    _stream = None

    def __init__(self):
        self.lib = _lib.lib()
        self._counters: Dict[torch.device, torch.Tensor] = {}

    # -- plumbing ---------------------------------------------------------------------------------------------------
    def check_tensor(self, t: torch.Tensor, what: str) -> None:
        _lib.require_device(t, what)

    def begin(self, device: torch.device) -> None:
        """Start of one pass (a few hundred calls): the current stream is looked up once."""
        self._stream = _lib.stream_ptr(device) if device.type == "cuda" else None

    def end(self) -> None:
        self._stream = None

    def stream(self, t: torch.Tensor):
        return self._stream if self._stream is not None else _lib.stream_ptr(t.device)

    def call(self, name: str, *args) -> None:
        _lib.check(getattr(self.lib, self.prefix + name)(*args))

    # -- data movement / normalisation ------------------------------------------------------------------------------
    def im2col(self, w: Window3d, src: torch.Tensor, voxel_stride: int, channel_stride: int, col: torch.Tensor) -> None:
        self.call("im2col", C.byref(w), src.data_ptr(), _DT[src.dtype], voxel_stride, channel_stride, col.data_ptr(),
                  _DT[col.dtype], col.stride(0), self.stream(col))

    def col2im(self, w: Window3d, dcol: torch.Tensor, dx: torch.Tensor) -> None:
        self.call("col2im", C.byref(w), dcol.data_ptr(), _DT[dcol.dtype], dcol.stride(0), dx.data_ptr(), dx.stride(0),
                  self.stream(dx))

    def pack_weight(self, weight: torch.Tensor, out: torch.Tensor) -> None:
        c_out, c_in = weight.shape[0], weight.shape[1]
        k3 = weight[0, 0].numel()
        self.call("pack_weight", weight.data_ptr(), c_out, c_in, k3, out.data_ptr(), _DT[out.dtype], out.stride(0),
                  self.stream(out))

    def unpack_wgrad(self, dw_mat: torch.Tensor, dweight: torch.Tensor) -> None:
        """dw_mat (c_out, ld) or (S, c_out, ld): S split-K partial results, summed on the way."""
        c_out, c_in = dweight.shape[0], dweight.shape[1]
        parts, pstride = (1, 0) if dw_mat.dim() == 2 else (dw_mat.shape[0], dw_mat.stride(0))
        self.call("unpack_wgrad", dw_mat.data_ptr(), dw_mat.stride(-2), c_out, c_in, dweight[0, 0].numel(), parts, pstride,
                  dweight.data_ptr(), self.stream(dweight))

    def bn_stats(self, x: torch.Tensor, sums: torch.Tensor) -> None:
        self.call("bn_stats", x.data_ptr(), _DT[x.dtype], x.shape[0], x.shape[1], x.stride(0), sums.data_ptr(),
                  self.stream(x))

    def bn_finalize(self, sums, V, Cn, eps, momentum, training, running_mean, running_var, mean, invstd) -> None:
        self.call("bn_finalize", _lib.ptr(sums), V, Cn, eps, momentum, int(training), _lib.ptr(running_mean),
                  _lib.ptr(running_var), mean.data_ptr(), _lib.ptr(invstd), self.stream(mean))

    def bn_apply(self, x, mean, invstd, gamma, beta, residual, relu, y) -> None:
        self.call("bn_apply", x.data_ptr(), _DT[x.dtype], x.shape[0], x.shape[1], x.stride(0), mean.data_ptr(),
                  invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _lib.ptr(residual),
                  0 if residual is None else residual.stride(0), int(relu), y.data_ptr(), y.stride(0), self.stream(y))

    def bn_backward_reduce(self, dy, dy2, y, x, mean, invstd, g_out, sums) -> None:
        self.call("bn_backward_reduce", dy.data_ptr(), _lib.ptr(dy2), _lib.ptr(y), x.data_ptr(), _DT[x.dtype],
                  x.shape[0], x.shape[1], x.stride(0), mean.data_ptr(), invstd.data_ptr(), g_out.data_ptr(),
                  sums.data_ptr(), self.stream(x))

    def bn_backward_apply(self, g, x, mean, invstd, gamma, sums, training, dx, dgamma, dbeta) -> None:
        self.call("bn_backward_apply", g.data_ptr(), x.data_ptr(), _DT[x.dtype], x.shape[0], x.shape[1], x.stride(0),
                  mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(), sums.data_ptr(), int(training), dx.data_ptr(),
                  _lib.ptr(dgamma), _lib.ptr(dbeta), self.stream(x))

    def maxpool(self, w: Window3d, x, y, argmax) -> None:
        self.call("maxpool", C.byref(w), x.data_ptr(), _DT[x.dtype], x.stride(0), y.data_ptr(), y.stride(0),
                  argmax.data_ptr(), self.stream(x))

    def maxpool_backward(self, w: Window3d, dy, dy2, argmax, dx) -> None:
        self.call("maxpool_backward", C.byref(w), dy.data_ptr(), _lib.ptr(dy2), _DT[dy.dtype], dy.stride(0), argmax.data_ptr(),
                  dx.data_ptr(), dx.stride(0), self.stream(dy))

    def broadcast_rows(self, v, scale: float, out) -> None:
        self.call("broadcast_rows", v.data_ptr(), scale, out.shape[0], out.shape[1], out.data_ptr(), _DT[out.dtype],
                  out.stride(0), self.stream(out))

    # -- contractions ---------------------------------------------------------------------------------------------------
    # A GEMM is described as (M, N, K, A, B, a_mn, b_mn, out):
    #   nt (0, 0): out(M, N) = A(M, K) B(N, K)^T   forward         (gathered input x packed weight)
    #   nn (0, 1): out(M, N) = A(M, K) B(K, N)     data gradient   (dY x packed weight, the weight read MN-major)
    #   tn (1, 1): out(M, N) = A(K, M)^T B(K, N)   weight gradient (contracts over the voxels, both operands MN-major)
    # bf16 operands: every description of a call goes out in ONE launch of the persistent tcgen05 job-list kernel
    # (independent GEMMs share the SMs); fp32 operands: one CUDA-core GEMM each (parity path).
    @staticmethod
    def _job(M, N, K, A, B, a_mn, b_mn, out) -> "_lib.GemmJob":
        j = _lib.GemmJob()
        j.M, j.N, j.K, j.A, j.lda, j.B, j.ldb = M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0)
        j.a_mn, j.b_mn, j.wait_job = a_mn, b_mn, -1
        j.b_static = 1
        j.bn = 256 if N >= 256 else (128 if (N >= 128 or b_mn) else 64)
        if out.dtype == torch.bfloat16:
            j.epi.out_bf16, j.epi.ld_bf16 = out.data_ptr(), out.stride(0)
        else:
            j.epi.out_f32, j.epi.ld_f32 = out.data_ptr(), out.stride(0)
        return j

    def run_gemms(self, descs) -> None:
        like = descs[0][7]
        if descs[0][3].dtype != torch.bfloat16:
            for (M, N, K, A, B, a_mn, b_mn, out) in descs:
                a_rs, a_cs = (1, A.stride(0)) if a_mn else (A.stride(0), 1)
                b_rs, b_cs = (1, B.stride(0)) if b_mn else (B.stride(0), 1)
                _lib.check(self.lib.neraf_gemm_f32(M, N, K, A.data_ptr(), a_rs, a_cs, B.data_ptr(), b_rs, b_cs, None, 0, None,
                                                   0, out.data_ptr(), out.stride(0), 0, self.stream(out)))
            return
        dev = like.device
        jobs = [self._job(*d) for d in descs]
        n_cnt = sum((int(j.M) + 255) // 256 for j in jobs) + 8
        cnt = self._counters.get(dev)
        if cnt is None or cnt.numel() < n_cnt:
            cnt = torch.zeros(max(n_cnt, 4096), dtype=torch.int32, device=dev)
            self._counters[dev] = cnt
        arr = (_lib.GemmJob * len(jobs))(*jobs)
        _lib.check(self.lib.neraf_gemm_bf16_jobs(arr, len(jobs), cnt.data_ptr(), cnt.numel() * 4, self.stream(like)))

    def gemm_nt(self, A, B, M, N, K, out) -> None:
        self.run_gemms([(M, N, K, A, B, 0, 0, out)])

    def gemm_nn(self, A, B, M, N, K, out) -> None:
        self.run_gemms([(M, N, K, A, B, 0, 1, out)])

    def gemm_tn(self, A, B, M, N, K, out) -> None:
        self.run_gemms([(M, N, K, A, B, 1, 1, out)])

    batched_backward = not PLAIN_FORMS      # False: one GEMM per launch, no split-K
    MAX_SPLITS = 16                         # the job list holds 24 GEMMs
    TILE_TARGET = 74                        # CTA pairs of a B200: enough weight-gradient tiles to occupy them

    def wgrad_splits(self, dy: torch.Tensor, v_out: int, kc: int, c_out: int) -> int:
        """Split-K factor of a weight-gradient GEMM: its output (c_out, kc) is a handful of 256-row tiles while the
        contraction runs over all voxels (the stem at 128^3: 4 tiles x 262 144), so the voxels are cut into chunks of
        >= 2048 that run as separate jobs of the same launch and are summed by unpack_wgrad."""
        if not self.batched_backward or dy.dtype != torch.bfloat16:
            return 1
        tiles = ((c_out + 255) // 256) * ((kc + 255) // 256)
        return max(1, min(self.MAX_SPLITS, self.TILE_TARGET // tiles, v_out // 2048))

    def gemm_backward(self, dy, wmat, col, v_out, kc, c_out, dw_parts, dcol) -> None:
        """Both gradients of one convolution from dy (v_out, c_out): the weight gradient dy^T col, split over the voxels
        into dw_parts.shape[0] partial matrices dw_parts (S, c_out, ld), and -- when dcol is given -- the data gradient
        dcol (v_out, kc) = dy wmat.  bf16: all of them are jobs of ONE launch."""
        S = dw_parts.shape[0]
        rows = (-(-v_out // S) + 63) // 64 * 64
        descs = []
        for s_ in range(S):
            r0, r1 = s_ * rows, min(v_out, (s_ + 1) * rows)
            descs.append((c_out, kc, r1 - r0, dy[r0:r1], col[r0:r1], 1, 1, dw_parts[s_]))
        if dcol is not None:
            descs.append((v_out, kc, c_out, dy, wmat, 0, 1, dcol))
        if self.batched_backward:
            self.run_gemms(descs)
        else:
            for d in descs:
                self.run_gemms([d])


_default_ops: Optional[GridOps] = None


def default_ops() -> GridOps:
    global _default_ops
    if _default_ops is None:
        _default_ops = GridOps()
    return _default_ops


# --------------------------------------------------------------------------------------------------------------------
# parameter holders with the reference's names and shapes
# --------------------------------------------------------------------------------------------------------------------
class Conv3dWeight(nn.Module):
    """The parameter of ``nn.Conv3d(c_in, c_out, k, stride, pad, bias=False)`` (NeRAF_resnet3d.py:26-29,119); xavier-normal
    initialised like the reference (:162-163)."""

    def __init__(self, c_in: int, c_out: int, k: int, stride: int = 1, pad: int = 0):
        super().__init__()
        self.c_in, self.c_out, self.k, self.stride, self.pad = c_in, c_out, k, stride, pad
        self.weight = nn.Parameter(torch.empty(c_out, c_in, k, k, k))
        nn.init.xavier_normal_(self.weight)


class BatchNorm3dParams(nn.Module):
    """Parameters and buffers of ``nn.BatchNorm3d`` (weight 1, bias 0: NeRAF_resnet3d.py:164-166)."""

    def __init__(self, channels: int, eps: float = 1e-5, momentum: float = 0.1):
        super().__init__()
        self.eps, self.momentum = eps, momentum
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))
        self.register_buffer("running_mean", torch.zeros(channels))
        self.register_buffer("running_var", torch.ones(channels))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


def _downsample(c_in: int, c_out: int, stride: int) -> nn.Sequential:
    return nn.Sequential(Conv3dWeight(c_in, c_out, 1, stride, 0), BatchNorm3dParams(c_out))   # NeRAF_resnet3d.py:171-175


class BasicBlock(nn.Module):      # NeRAF_resnet3d.py:44-77
    expansion = 1

    def __init__(self, in_planes: int, planes: int, stride: int = 1, downsample: Optional[nn.Module] = None):
        super().__init__()
        self.conv1 = Conv3dWeight(in_planes, planes, 3, stride, 1)
        self.bn1 = BatchNorm3dParams(planes)
        self.conv2 = Conv3dWeight(planes, planes, 3, 1, 1)
        self.bn2 = BatchNorm3dParams(planes)
        self.downsample = downsample
        self.stride = stride

    def units(self):
        return [(self.conv1, self.bn1), (self.conv2, self.bn2)]


class Bottleneck(nn.Module):      # NeRAF_resnet3d.py:80-122
    expansion = 4

    def __init__(self, in_planes: int, planes: int, stride: int = 1, downsample: Optional[nn.Module] = None):
        super().__init__()
        self.conv1 = Conv3dWeight(in_planes, planes, 1)
        self.bn1 = BatchNorm3dParams(planes)
        self.conv2 = Conv3dWeight(planes, planes, 3, stride, 1)
        self.bn2 = BatchNorm3dParams(planes)
        self.conv3 = Conv3dWeight(planes, planes * 4, 1)
        self.bn3 = BatchNorm3dParams(planes * 4)
        self.downsample = downsample
        self.stride = stride

    def units(self):
        return [(self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)]


# --------------------------------------------------------------------------------------------------------------------
# execution
# --------------------------------------------------------------------------------------------------------------------
class _Act:
    """A channels-last activation: matrix (V, C) + its spatial extent."""
    __slots__ = ("t", "dims")

    def __init__(self, t: torch.Tensor, dims: Tuple[int, int, int]):
        self.t, self.dims = t, dims


class _UnitRecord:
    """What the backward pass of one conv + batch-norm unit needs."""
    __slots__ = ("conv", "bn", "window", "col", "wmat", "xc", "y", "mean", "invstd", "relu", "in_dims", "c_in", "direct")


class _Runner:
    """One forward (and, from the records it leaves, one backward) pass of the network through ``GridOps``."""

    def __init__(self, ops: GridOps, dtype: torch.dtype, training: bool, keep: bool):
        self.ops, self.dtype, self.training, self.keep = ops, dtype, training, keep
        self.grads: Dict[int, torch.Tensor] = {}
        self.tracked: List[torch.Tensor] = []

    # ---- forward ------------------------------------------------------------------------------------------------------
    def conv_bn(self, conv: Conv3dWeight, bn: BatchNorm3dParams, x: _Act, relu: bool, residual: Optional[_Act] = None,
                grid: Optional[torch.Tensor] = None) -> Tuple[_Act, _UnitRecord]:
        ops, dev = self.ops, conv.weight.device
        weight = conv.weight.detach()
        if grid is not None and self.dtype == torch.bfloat16 and not PLAIN_FORMS:
            # The stem reads the reference's channels-first fp32 grid (1, C, D, H, W).  C = 7 would force the 2-byte
            # gather on the largest matrix of the network (0.46 GB at 128^3), so the grid is first laid out channels-last
            # with the channels padded to 8 (a k = 1 gather: 33 MB) and the stem runs on that -- 16-byte gather, K = 8 k^3
            # with zero weights on the pad channel.
            if grid.shape[1] != conv.c_in:
                raise ValueError(f"convolution expects {conv.c_in} input channels, got {grid.shape[1]}")
            dims = tuple(grid.shape[2:])
            c_pad = _round_up(grid.shape[1], 8)
            v_in = dims[0] * dims[1] * dims[2]
            cl = torch.empty(v_in, c_pad, dtype=self.dtype, device=dev)
            ops.im2col(Window3d(dims[0], dims[1], dims[2], grid.shape[1], 1, 1, 0), grid, 1, v_in, cl)
            x, grid = _Act(cl, dims), None
            if c_pad != conv.c_in:
                weight = torch.nn.functional.pad(weight, (0, 0, 0, 0, 0, 0, 0, c_pad - conv.c_in))
        if grid is not None:                       # fp32: the stem gathers straight from the (1, C, D, H, W) grid
            in_dims, c_in = tuple(grid.shape[2:]), grid.shape[1]
        else:
            in_dims, c_in = x.dims, x.t.shape[1]
        if c_in != weight.shape[1]:
            raise ValueError(f"convolution expects {conv.c_in} input channels, got {c_in}")
        w = Window3d(in_dims[0], in_dims[1], in_dims[2], c_in, conv.k, conv.stride, conv.pad)
        out_dims = w.out_dims
        v_out = out_dims[0] * out_dims[1] * out_dims[2]
        kc = conv.k ** 3 * c_in
        direct = grid is None and conv.k == 1 and conv.stride == 1 and x.t.stride(0) % 8 == 0
        if direct:
            col = x.t                                                                  # a 1x1x1 convolution is a GEMM
        else:
            col = torch.empty(v_out, _round_up(kc, 8), dtype=self.dtype, device=dev)
            if grid is not None:
                ops.im2col(w, grid, 1, in_dims[0] * in_dims[1] * in_dims[2], col)
            else:
                ops.im2col(w, x.t, x.t.stride(0), 1, col)
        wmat = torch.empty(conv.c_out, _round_up(kc, 8), dtype=self.dtype, device=dev)
        ops.pack_weight(weight, wmat)
        xc = torch.empty(v_out, conv.c_out, dtype=self.dtype, device=dev)
        ops.gemm_nt(col, wmat, v_out, conv.c_out, kc, xc)

        mean = torch.empty(conv.c_out, dtype=torch.float32, device=dev)
        invstd = torch.empty_like(mean)
        if self.training:
            sums = torch.empty(2, conv.c_out, dtype=torch.float64, device=dev)
            ops.bn_stats(xc, sums)
            ops.bn_finalize(sums, v_out, conv.c_out, bn.eps, bn.momentum, True, bn.running_mean, bn.running_var, mean, invstd)
            self.tracked.append(bn.num_batches_tracked)          # all counters are bumped by ONE call after the pass
        else:
            ops.bn_finalize(None, v_out, conv.c_out, bn.eps, 0.0, False, bn.running_mean, bn.running_var, mean, invstd)
        y = torch.empty_like(xc)
        ops.bn_apply(xc, mean, invstd, bn.weight.detach(), bn.bias.detach(), None if residual is None else residual.t,
                     relu, y)
        rec = None
        if self.keep:
            rec = _UnitRecord()
            rec.conv, rec.bn, rec.window, rec.col, rec.wmat, rec.xc, rec.y = conv, bn, w, col, wmat, xc, y
            rec.mean, rec.invstd, rec.relu, rec.in_dims, rec.c_in, rec.direct = mean, invstd, relu, in_dims, c_in, direct
        return _Act(y, out_dims), rec

    # ---- backward -----------------------------------------------------------------------------------------------------
    def conv_bn_backward(self, rec: _UnitRecord, dy: torch.Tensor, dy2: Optional[torch.Tensor], keep_g: bool,
                         need_dx: bool) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
        """Returns (gradient w.r.t. the unit's input or None, g = the gradient that reached the normalisation)."""
        ops, conv, bn = self.ops, rec.conv, rec.bn
        dev = rec.xc.device
        v_out, c_out = rec.xc.shape
        sums = torch.empty(2, c_out, dtype=torch.float64, device=dev)
        g = torch.empty_like(rec.xc)
        ops.bn_backward_reduce(dy, dy2, rec.y if rec.relu else None, rec.xc, rec.mean, rec.invstd, g, sums)
        dxc = torch.empty_like(g) if keep_g else g
        dgamma = torch.empty(c_out, dtype=torch.float32, device=dev)
        dbeta = torch.empty_like(dgamma)
        ops.bn_backward_apply(g, rec.xc, rec.mean, rec.invstd, bn.weight.detach(), sums, self.training, dxc, dgamma, dbeta)
        self.grads[id(bn.weight)] = dgamma
        self.grads[id(bn.bias)] = dbeta

        c_in = rec.c_in                                      # the stem's bf16 path runs on channels padded to 8
        kc = conv.k ** 3 * c_in
        n_split = ops.wgrad_splits(dxc, v_out, kc, c_out)
        dw_parts = torch.empty(n_split, c_out, rec.wmat.stride(0), dtype=torch.float32, device=dev)
        dcol = None
        if need_dx:                  # a 1x1x1 stride-1 convolution's "gathered" gradient IS the input gradient
            dcol = torch.empty(v_out, c_in if rec.direct else rec.wmat.stride(0), dtype=self.dtype, device=dev)
        ops.gemm_backward(dxc, rec.wmat, rec.col, v_out, kc, c_out, dw_parts, dcol)
        if n_split == 1 and conv.k == 1 and dw_parts.stride(1) == kc and c_in == conv.c_in:
            dweight = dw_parts.view(conv.weight.shape)
        else:
            dweight = torch.empty(c_out, c_in, conv.k, conv.k, conv.k, dtype=torch.float32, device=dev)
            ops.unpack_wgrad(dw_parts, dweight)
            if c_in != conv.c_in:
                dweight = dweight[:, :conv.c_in].contiguous()
        self.grads[id(conv.weight)] = dweight

        dx = None
        if need_dx:
            if rec.direct:
                dx = dcol
            else:
                v_in = rec.in_dims[0] * rec.in_dims[1] * rec.in_dims[2]
                dx = torch.empty(v_in, c_in, dtype=self.dtype, device=dev)
                ops.col2im(rec.window, dcol, dx)
        return dx, g


class _GridNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net: "ResNet3D", x: torch.Tensor, *params: torch.Tensor):
        ops = net.ops if net.ops is not None else default_ops()
        ops.check_tensor(x, "grid")
        ops.check_tensor(params[0], "ResNet3D parameters")
        need_grad = any(ctx.needs_input_grad[2:])
        ops.begin(params[0].device)
        try:
            with torch.no_grad():
                out, tape = net._run_forward(ops, x, keep=need_grad)
        finally:
            ops.end()
        ctx.net, ctx.ops, ctx.tape, ctx.params = net, ops, tape, params
        return out

    @staticmethod
    def backward(ctx, dout: torch.Tensor):
        ctx.ops.begin(dout.device)
        try:
            with torch.no_grad():
                grads = ctx.net._run_backward(ctx.ops, ctx.tape, dout)
        finally:
            ctx.ops.end()
        ctx.tape = None
        return (None, None) + tuple(grads.get(id(p)) for p in ctx.params)


class ResNet3D(nn.Module):
    """NeRAF_resnet3d.py:116-201 (forward :178-193).  ``precision``: "bf16" (tcgen05) or "fp32" (parity path)."""

    def __init__(self, in_channels: int, block, layers: Sequence[int], grid_step: Optional[float] = None,
                 N_features: int = 1024, precision: str = "bf16"):
        super().__init__()
        assert N_features in [1024, 2048], 'N_features should be 1024 or 2048'
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        self.precision = precision
        self.ops: Optional[GridOps] = None
        self.in_planes = 64
        self.conv1 = Conv3dWeight(in_channels, 64, 5, 2, 2)            # 128 -> 64
        self.bn1 = BatchNorm3dParams(64)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.N_features = N_features
        if N_features == 2048:
            self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        if grid_step is None:
            grid_step = 1 / 128
        # the reference's nn.AvgPool3d window (:141-157); the feature is (1, N, 1, 1, 1) when the last extent equals it
        if grid_step >= 1 / 64 - 1 / 512:
            self.avgpool_size = 2 if N_features == 2048 else 4
        elif grid_step >= 1 / 128 - 1 / 512:
            self.avgpool_size = 4 if N_features == 2048 else 8
        else:
            self.avgpool_size = 8 if N_features == 2048 else 16

    def _make_layer(self, block, planes: int, blocks: int, stride: int = 1) -> nn.Sequential:
        downsample = None
        if stride != 1 or self.in_planes != planes * block.expansion:
            downsample = _downsample(self.in_planes, planes * block.expansion, stride)
        layers = [block(self.in_planes, planes, stride, downsample)]
        self.in_planes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.in_planes, planes))
        return nn.Sequential(*layers)

    def _stages(self) -> List[nn.Sequential]:
        stages = [self.layer1, self.layer2, self.layer3]
        if self.N_features == 2048:
            stages.append(self.layer4)
        return stages

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 5 or x.shape[0] != 1:
            raise ValueError(f"the grid must be (1, C, D, H, W) like the reference's feat_grid, got {tuple(x.shape)}")
        params = self.__dict__.get("_param_list")
        if params is None:                      # walked once: Module.to / load_state_dict keep the Parameter objects
            params = [p for p in self.parameters()]
            self.__dict__["_param_list"] = params
        return _GridNetFn.apply(self, x, *params)

    # ---- the network as a sequence of library calls -----------------------------------------------------------------------
    def _run_forward(self, ops: GridOps, x: torch.Tensor, keep: bool):
        dtype = torch.bfloat16 if self.precision == "bf16" else torch.float32
        run = _Runner(ops, dtype, self.training, keep)
        grid = x.detach()
        if grid.dtype != torch.float32 or not grid.is_contiguous():
            grid = grid.float().contiguous()
        dev = grid.device
        tape: Dict = {"run": run, "blocks": []}
        a, tape["stem"] = run.conv_bn(self.conv1, self.bn1, None, relu=True, grid=grid)        # :179-181
        wp = Window3d(a.dims[0], a.dims[1], a.dims[2], a.t.shape[1], 3, 2, 1)                  # :122,182
        pd = wp.out_dims
        pooled = torch.empty(pd[0] * pd[1] * pd[2], a.t.shape[1], dtype=dtype, device=dev)
        argmax = torch.empty(pooled.shape, dtype=torch.int32, device=dev)
        ops.maxpool(wp, a.t, pooled, argmax)
        tape["pool"] = (wp, argmax, a.t.shape)
        a = _Act(pooled, pd)
        for stage in self._stages():
            for block in stage:
                block_in = a
                units = block.units()
                recs = []
                shortcut, ds_rec = block_in, None
                if block.downsample is not None:                                               # :105-106
                    shortcut, ds_rec = run.conv_bn(block.downsample[0], block.downsample[1], block_in, relu=False)
                for i, (conv, bn) in enumerate(units):
                    last = i == len(units) - 1
                    a, rec = run.conv_bn(conv, bn, a, relu=True, residual=shortcut if last else None)
                    recs.append(rec)
                tape["blocks"].append((recs, ds_rec))
        v, c = a.t.shape
        if a.dims != (self.avgpool_size,) * 3:
            raise _lib.NerafError(f"the last stage is {a.dims}, the average pooling window is {self.avgpool_size}^3: only the "
                                  "reference's default case (one pooled voxel) is implemented")
        sums = torch.empty(2, c, dtype=torch.float64, device=dev)
        feat = torch.empty(c, dtype=torch.float32, device=dev)
        ops.bn_stats(a.t, sums)                                                                 # global average pooling
        ops.bn_finalize(sums, v, c, 0.0, 0.0, True, None, None, feat, None)
        tape["last_shape"] = (v, c)
        if run.tracked:
            torch._foreach_add_(run.tracked, 1)
        return feat.view(1, c, 1, 1, 1), (tape if keep else None)

    def _run_backward(self, ops: GridOps, tape: Dict, dout: torch.Tensor) -> Dict[int, torch.Tensor]:
        run: _Runner = tape["run"]
        v, c = tape["last_shape"]
        dev = dout.device
        d = torch.empty(v, c, dtype=run.dtype, device=dev)
        ops.broadcast_rows(dout.detach().float().contiguous().view(-1), 1.0 / v, d)
        dy, dy2 = d, None                                    # gradient w.r.t. the current block's output: dy (+ dy2)
        for recs, ds_rec in reversed(tape["blocks"]):
            n = len(recs)
            g_last = None
            for i in range(n - 1, -1, -1):
                last = i == n - 1
                dx, g = run.conv_bn_backward(recs[i], dy, dy2 if last else None, keep_g=last, need_dx=True)
                if last:
                    g_last = g
                dy, dy2 = dx, None
            d_main = dy
            if ds_rec is not None:
                d_short, _ = run.conv_bn_backward(ds_rec, g_last, None, keep_g=False, need_dx=True)
            else:
                d_short = g_last
            dy, dy2 = d_main, d_short
        wp, argmax, stem_shape = tape["pool"]
        d_stem = torch.empty(stem_shape, dtype=run.dtype, device=dev)
        ops.maxpool_backward(wp, dy, dy2, argmax, d_stem)       # dy + dy2: the first block's main path and its shortcut
        run.conv_bn_backward(tape["stem"], d_stem, None, keep_g=False, need_dx=False)
        return run.grads


def conv_flops(net: ResNet3D, n: int) -> float:
    """Algorithmic FLOPs of one training step (forward + data gradient + weight gradient of every convolution; the stem
    has no data gradient) on an n^3 grid -- the contraction work, 2 FLOP per multiply-add."""
    def out(e, conv):
        return (e + 2 * conv.pad - conv.k) // conv.stride + 1

    e = out(n, net.conv1)
    total = 2.0 * e ** 3 * net.conv1.k ** 3 * net.conv1.c_in * net.conv1.c_out * 2           # forward + wgrad
    e = (e + 2 - 3) // 2 + 1                                                                  # max pooling
    for stage in net._stages():
        for block in stage:
            e_in = e
            for conv, _ in block.units():
                e = out(e, conv)
                total += 2.0 * e ** 3 * conv.k ** 3 * conv.c_in * conv.c_out * 3
            if block.downsample is not None:
                ds = block.downsample[0]
                total += 2.0 * out(e_in, ds) ** 3 * ds.c_in * ds.c_out * 3
    return total


# NeRAF_resnet3d.py:204-262
def resnet18(in_channels=3, pretrained=False, grid_step=None, N_features=None, **kwargs):
    return ResNet3D(in_channels, BasicBlock, [2, 2, 2, 2], grid_step=grid_step, N_features=N_features, **kwargs)


def resnet34(in_channels=3, pretrained=False, grid_step=None, N_features=None, **kwargs):
    return ResNet3D(in_channels, BasicBlock, [3, 4, 6, 3], grid_step=grid_step, N_features=N_features, **kwargs)


def resnet50(in_channels=3, pretrained=False, grid_step=None, N_features=None, **kwargs):
    if pretrained:
        raise _lib.NerafError("pretrained=True downloads 2-D ImageNet weights in the reference (and cannot load them "
                              "into 3-D convolutions); NeRAF always passes pretrained=False (NeRAF_model.py:185)")
    return ResNet3D(in_channels, Bottleneck, [3, 4, 6, 3], grid_step=grid_step, N_features=N_features, **kwargs)


def resnet101(in_channels=3, pretrained=False, grid_step=None, N_features=None, **kwargs):
    return ResNet3D(in_channels, Bottleneck, [3, 4, 23, 3], grid_step=grid_step, N_features=N_features, **kwargs)


def resnet152(in_channels=3, pretrained=False, grid_step=None, N_features=None, **kwargs):
    return ResNet3D(in_channels, Bottleneck, [3, 8, 36, 3], grid_step=grid_step, N_features=N_features, **kwargs)


class ResNet3D_helper(nn.Module):
    """NeRAF_resnet3d.py:265-299, constructed at NeRAF_model.py:185 as
    ``ResNet3D_helper(in_channels=7, backbone='resnet50', pretrained=False, grid_step=..., N_features=...)``."""

    backbones = {"resnet18": resnet18, "resnet34": resnet34, "resnet50": resnet50, "resnet101": resnet101,
                 "resnet152": resnet152}

    def __init__(self, in_channels=3, backbone="resnet50", pretrained=False, grid_step=None, N_features=1024,
                 precision: str = "bf16"):
        super().__init__()
        self.backbone_net = ResNet3D_helper.backbones[backbone](in_channels=in_channels, pretrained=pretrained,
                                                                grid_step=grid_step, N_features=N_features,
                                                                precision=precision)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.backbone_net(x)
