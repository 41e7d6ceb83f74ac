"""ctypes binding of the C ABI in include/neraf_b200.h.

This is the only place the shared library is opened.  There is deliberately no
fallback: if ``neraf_b200/lib/libneraf_b200.so`` is missing or no sm_100 device
is present, every product entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional, Sequence

import torch

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libneraf_b200.so")

PREC_FP32, PREC_BF16 = 0, 1
ACT_NONE, ACT_LEAKY, ACT_TANH10 = 0, 1, 2
CRIT_SC_SLMSE, CRIT_SC_SLL1, CRIT_MSE = 0, 1, 2
ORDER_TIME_MIC_SRC_ROT, ORDER_MIC_SRC_TIME_ROT = 0, 1
MAX_TRUNK = 8
MAX_RANKS, EXCHANGE_BYTES, MAX_EXCHANGE_CHUNKS, NOTIFY_COUNTERS = 16, 8192, 32, 1024
EXCHANGE_STATE_WORDS = 64

PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16}
CRITERIA = {"SC+SLMSE": CRIT_SC_SLMSE, "SC+SLL1": CRIT_SC_SLL1, "MSE": CRIT_MSE}


class NerafError(RuntimeError):
    """Non-zero status from the CUDA library (text from neraf_last_error())."""


class FieldDims(C.Structure):
    _fields_ = [("n_grid", C.c_int32), ("n_enc", C.c_int32), ("n_trunk", C.c_int32),
                ("trunk", C.c_int32 * MAX_TRUNK), ("n_channels", C.c_int32), ("n_freq", C.c_int32)]


class Queries(C.Structure):
    _fields_ = [("batch", C.c_int64), ("time_query", C.c_void_p), ("mic_pose", C.c_void_p),
                ("source_pose", C.c_void_p), ("rot", C.c_void_p), ("aabb", C.c_void_p),
                ("time_denominator", C.c_float), ("order", C.c_int32), ("enc", C.c_void_p), ("enc_ld", C.c_int64)]


class GemmEpilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("act", C.c_int32), ("gate", C.c_void_p), ("ldg", C.c_int64),
                ("out_bf16", C.c_void_p), ("ld_bf16", C.c_int64), ("out_bf16_t", C.c_void_p), ("ld_t", C.c_int64),
                ("out_f32", C.c_void_p), ("ld_f32", C.c_int64), ("accumulate_f32", C.c_int32),
                ("mask_out", C.c_void_p), ("gate_mask", C.c_void_p), ("ld_mask", C.c_int64),
                ("out_f32_multicast", C.c_void_p), ("loss_gt", C.c_void_p), ("ld_gt", C.c_int64),
                ("loss_sums", C.c_void_p)]


class GemmJob(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64), ("A", C.c_void_p), ("lda", C.c_int64),
                ("B", C.c_void_p), ("ldb", C.c_int64), ("a_mn", C.c_int32), ("b_mn", C.c_int32), ("b_static", C.c_int32),
                ("bn", C.c_int32),
                ("wait_job", C.c_int32), ("wait_all", C.c_int32), ("merge_next", C.c_int32), ("epi", GemmEpilogue), ("colsum", C.c_void_p),
                ("notify", C.c_void_p)]


class Multicast(C.Structure):
    _fields_ = [("local_base", C.c_void_p), ("multicast_base", C.c_void_p), ("bytes", C.c_size_t)]


class RankExchange(C.Structure):    # neraf_rank_exchange
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("peers", C.c_void_p * MAX_RANKS)]


class LossGrad(C.Structure):        # neraf_loss_grad
    _fields_ = [("gt", C.c_void_p), ("n_total", C.c_int64), ("criterion", C.c_int32), ("sums", C.c_void_p),
                ("w_sc", C.c_float), ("w_mag", C.c_float), ("losses", C.c_void_p), ("total", C.c_void_p),
                ("fuse_sums", C.c_int32), ("sync", C.c_void_p), ("exchange", C.POINTER(RankExchange))]


class ExchangeChunk(C.Structure):   # neraf_exchange_chunk
    _fields_ = [("offset", C.c_int64), ("bytes", C.c_int64), ("notify", C.c_void_p), ("notify_count", C.c_uint32),
                ("notify_increment", C.c_uint32), ("f32", C.c_int32), ("dst", C.c_void_p), ("dst_ld", C.c_int64),
                ("row_elems", C.c_int32), ("src_ld", C.c_int32)]


class GradExchange(C.Structure):    # neraf_grad_exchange
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("n_chunks", C.c_int32), ("max_ctas", C.c_int32),
                ("chunks", ExchangeChunk * MAX_EXCHANGE_CHUNKS), ("multicast", C.c_void_p),
                ("peers", C.c_void_p * MAX_RANKS), ("signals", C.c_void_p * MAX_RANKS), ("state", C.c_void_p),
                ("pull", C.c_int32), ("trace", C.c_void_p)]


class DpOptions(C.Structure):
    _fields_ = [("mc", C.POINTER(Multicast)), ("dw0_compact", C.c_void_p), ("defer_grid_grads", C.c_int32),
                ("phase", C.c_int32), ("max_ctas", C.c_int32), ("loss", C.POINTER(LossGrad)),
                ("dweights_bf16", C.POINTER(C.c_void_p)), ("notify", C.c_void_p),
                ("notify_offset", C.POINTER(C.c_uint32)), ("notify_count", C.POINTER(C.c_uint32)),
                ("notify_increment", C.POINTER(C.c_uint32)), ("exchange", C.POINTER(GradExchange)),
                ("zero_tail_slack", C.c_int32)]


class MetricParams(C.Structure):     # neraf_metric_params
    _fields_ = [("n_samples", C.c_int32), ("fs", C.c_double), ("t60_decay_db", C.c_float), ("t60_highpass_hz", C.c_double)]


class Window3d(C.Structure):         # neraf_window3d
    _fields_ = [("in_d", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("channels", C.c_int32),
                ("k", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32)]


class GlParams(C.Structure):
    _fields_ = [("n_fft", C.c_int32), ("win_length", C.c_int32), ("hop", C.c_int32), ("n_frames", C.c_int32),
                ("n_iter", C.c_int32), ("momentum", C.c_float), ("input_is_log", C.c_int32)]


_vp, _i64, _i32, _f32, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t
_pp = C.POINTER(C.c_void_p)

_pw = C.POINTER(Window3d)
# the grid-feature producer's operators; tests/test_gridnet.py binds a host build of the same code with the same lists
GRID_SIGNATURES = {
    "neraf_grid_im2col": (C.c_int, [_pw, _vp, _i32, _i64, _i64, _vp, _i32, _i64, _vp]),
    "neraf_grid_col2im": (C.c_int, [_pw, _vp, _i32, _i64, _vp, _i64, _vp]),
    "neraf_grid_pack_weight": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _i32, _i64, _vp]),
    "neraf_grid_unpack_wgrad": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _i32, _i64, _vp, _vp]),
    "neraf_grid_bn_stats": (C.c_int, [_vp, _i32, _i64, _i64, _i64, _vp, _vp]),
    "neraf_grid_bn_finalize": (C.c_int, [_vp, _i64, _i64, _f32, _f32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "neraf_grid_bn_apply": (C.c_int, [_vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i64, _vp]),
    "neraf_grid_bn_backward_reduce": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "neraf_grid_bn_backward_apply": (C.c_int, [_vp, _vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp,
                                                _vp]),
    "neraf_grid_maxpool": (C.c_int, [_pw, _vp, _i32, _i64, _vp, _i64, _vp, _vp]),
    "neraf_grid_maxpool_backward": (C.c_int, [_pw, _vp, _vp, _i32, _i64, _vp, _vp, _i64, _vp]),
    "neraf_grid_broadcast_rows": (C.c_int, [_vp, _f32, _i64, _i64, _vp, _i32, _i64, _vp]),
}

# name -> (restype, argtypes); must list every symbol declared in include/neraf_b200.h
SIGNATURES = {
    "neraf_version": (C.c_int, []),
    "neraf_last_error": (C.c_char_p, []),
    "neraf_launch_count": (C.c_longlong, []),
    "neraf_abi_sizeof": (C.c_size_t, [C.c_char_p]),
    "neraf_device_supported": (C.c_int, []),
    "neraf_field_sizes": (C.c_int, [C.POINTER(FieldDims), _i32, _i64, C.POINTER(_sz), C.POINTER(_sz)]),
    "neraf_field_pack": (C.c_int, [C.POINTER(FieldDims), _i32, _pp, _pp, _vp, _sz, _vp]),
    "neraf_field_forward": (C.c_int, [C.POINTER(FieldDims), _i32, C.POINTER(Queries), _vp, _pp, _pp, _vp, _sz, _i32,
                                       _vp, _sz, _vp, _i32, _vp]),
    "neraf_field_forward_loss_sums": (C.c_int, [C.POINTER(FieldDims), _i32, C.POINTER(Queries), _vp, _pp, _pp, _vp, _sz,
                                                 _i32, _vp, _sz, _vp, _i32, _vp, _vp, _vp]),
    "neraf_field_backward": (C.c_int, [C.POINTER(FieldDims), _i32, _i64, _vp, _vp, _vp, _pp, _vp, _vp, _sz, _pp, _pp,
                                        _vp, _vp, _i64, _vp]),
    "neraf_field_backward_dp": (C.c_int, [C.POINTER(FieldDims), _i32, _i64, _vp, _vp, _vp, _pp, _vp, _vp, _sz, _pp, _pp,
                                           _vp, _vp, _i64, C.POINTER(DpOptions), _vp]),
    "neraf_dp_exchange_grads": (C.c_int, [C.POINTER(GradExchange), _vp]),
    "neraf_field_grid_grads": (C.c_int, [C.POINTER(FieldDims), _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "neraf_encode_queries": (C.c_int, [C.POINTER(Queries), _vp, _i64, _vp]),
    "neraf_spectral_loss_sums": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _vp]),
    "neraf_spectral_loss_finalize": (C.c_int, [_vp, _i64, _i32, _f32, _f32, _vp, _vp]),
    "neraf_acoustic_metrics": (C.c_int, [C.POINTER(MetricParams), _vp, _i64, _vp, _sz, _vp, _vp, _vp, _vp]),
    "neraf_gather_batch": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "neraf_spectral_loss_forward": (C.c_int, [_vp, _vp, _i64, _i32, _f32, _f32, _vp, _vp, _vp]),
    "neraf_spectral_loss_backward": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _f32, _f32, _vp, _vp]),
    "neraf_griffinlim_sizes": (C.c_int, [C.POINTER(GlParams), _i64, C.POINTER(_sz)]),
    "neraf_griffinlim": (C.c_int, [C.POINTER(GlParams), _i64, _i32, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _i64,
                                    _i64, _vp, _sz, _vp, _vp]),
    "neraf_gemm_f32": (C.c_int, [_i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i32, _vp, _i64, _vp, _i64,
                                  _i32, _vp]),
    "neraf_gemm_bf16": (C.c_int, [_i64, _i64, _i64, _vp, _i64, _vp, _i64, C.POINTER(GemmEpilogue), _vp]),
    "neraf_gemm_bf16_jobs": (C.c_int, [C.POINTER(GemmJob), C.c_int, _vp, _sz, _vp]),
    "neraf_gemm_bf16_set_tile": (C.c_int, [C.c_int]),
    "neraf_convert_bf16": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp]),
    **GRID_SIGNATURES,
}

# ctypes mirror of every public struct, by its C type name (layouts are checked against neraf_abi_sizeof)
STRUCTS = {"neraf_field_dims": FieldDims, "neraf_queries": Queries, "neraf_multicast": Multicast,
           "neraf_rank_exchange": RankExchange, "neraf_loss_grad": LossGrad, "neraf_dp_options": DpOptions,
           "neraf_exchange_chunk": ExchangeChunk, "neraf_grad_exchange": GradExchange, "neraf_gl_params": GlParams,
           "neraf_metric_params": MetricParams, "neraf_gemm_epilogue": GemmEpilogue, "neraf_gemm_job": GemmJob,
           "neraf_window3d": Window3d}

_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """The loaded shared library (opened once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise NerafError(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(or `make -C neraf_b200/csrc`). neraf_b200 has no CPU fallback.")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise NerafError(f"neraf_b200 error {rc}: {lib().neraf_last_error().decode(errors='replace')}")


def require_device(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise NerafError(f"{what} must live on a CUDA device (got {t.device}); neraf_b200 has no CPU path")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr(device: Optional[torch.device] = None) -> int:
    """cudaStream_t of torch's current stream on ``device``.  The raw getter is one C call; going through
    ``torch.cuda.current_stream(device).cuda_stream`` costs ~7 us of Python per call, four times per eager step."""
    if _raw_stream is not None:
        idx = device.index if (device is not None and device.index is not None) else torch.cuda.current_device()
        return _raw_stream(idx)
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def ptr_array(tensors: Sequence[torch.Tensor]):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def make_dims(n_grid: int, n_enc: int, trunk: Sequence[int], n_channels: int, n_freq: int) -> FieldDims:
    if len(trunk) > MAX_TRUNK:
        raise ValueError(f"at most {MAX_TRUNK} trunk layers")
    d = FieldDims()
    d.n_grid, d.n_enc, d.n_trunk, d.n_channels, d.n_freq = n_grid, n_enc, len(trunk), n_channels, n_freq
    for i, w in enumerate(trunk):
        d.trunk[i] = w
    return d
