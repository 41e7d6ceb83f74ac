"""GPU parity tests of the operator-level C-ABI entry points (GEMMs, encodings, loss) against
plain torch fp64 / the oracle on the same seeded inputs."""
import ctypes as C

import numpy as np
import pytest
import torch

from neraf_b200 import _lib
from neraf_b200 import synthetic as syn
from oracle import encodings as oenc
from oracle import loss as oloss
from tests.util import cuda, rel_fro, rel_max

pytestmark = pytest.mark.gpu


def _gemm_f32(A, a_rs, a_cs, B, b_rs, b_cs, M, N, K, bias=None, act=0, gate=None, Cinit=None, accumulate=0):
    lib = _lib.lib()
    out = torch.zeros(M, N, device=A.device) if Cinit is None else Cinit.clone()
    _lib.check(lib.neraf_gemm_f32(M, N, K, A.data_ptr(), a_rs, a_cs, B.data_ptr(), b_rs, b_cs, _lib.ptr(bias), act,
                                  _lib.ptr(gate), 0 if gate is None else gate.stride(0), out.data_ptr(), out.stride(0),
                                  accumulate, _lib.stream_ptr(A.device)))
    return out


def _act(v, act):
    if act == 1:
        return torch.nn.functional.leaky_relu(v, 0.1)
    if act == 2:
        return 10 * torch.tanh(v)
    return v


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (37, 65, 19), (128, 128, 16), (200, 513, 163), (257, 130, 300)])
def test_gemm_f32_all_layouts(M, N, K):
    dev = cuda()
    g = torch.Generator(device="cpu").manual_seed(M * 1000 + N)
    A = torch.randn(M, K, generator=g).to(dev)
    B = torch.randn(N, K, generator=g).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    gate = torch.randn(M, N, generator=g).to(dev)
    ref = A.double() @ B.double().t()
    # NT (both K-contiguous)
    out = _gemm_f32(A, K, 1, B, K, 1, M, N, K)
    assert rel_fro(out, ref) < 2e-6
    # transposed storage for both operands (the wgrad form)
    At, Bt = A.t().contiguous(), B.t().contiguous()
    out = _gemm_f32(At, 1, M, Bt, 1, N, M, N, K)
    assert rel_fro(out, ref) < 2e-6
    # mixed (the dgrad form) + full epilogue
    for act in (0, 1, 2):
        full = _act(ref + bias.double(), act) * torch.where(gate > 0, 1.0, 0.1).double()
        out = _gemm_f32(A, K, 1, Bt, 1, N, M, N, K, bias=bias, act=act, gate=gate)
        assert rel_fro(out, full) < 2e-6, act
    # accumulate happens on the pre-activation
    c0 = torch.randn(M, N, generator=g).to(dev)
    out = _gemm_f32(A, K, 1, B, K, 1, M, N, K, bias=bias, act=1, Cinit=c0, accumulate=1)
    assert rel_fro(out, _act(ref + c0.double() + bias.double(), 1)) < 2e-6


def _gemm_bf16(A, B, M, N, K, bias=None, act=0, gate=None, want=("rm", "t", "f32"), f32_init=None, accumulate=0):
    """A (M, lda) bf16, B (N, ldb) bf16 device tensors (K-major)."""
    lib = _lib.lib()
    dev = A.device
    e = _lib.GemmEpilogue()
    outs = {}
    ldn, ldm = (N + 7) // 8 * 8, (M + 7) // 8 * 8
    if "rm" in want:
        outs["rm"] = torch.full((M, ldn), 7.0, dtype=torch.bfloat16, device=dev)
        e.out_bf16, e.ld_bf16 = outs["rm"].data_ptr(), ldn
    if "t" in want:
        outs["t"] = torch.full((N, ldm), 7.0, dtype=torch.bfloat16, device=dev)
        e.out_bf16_t, e.ld_t = outs["t"].data_ptr(), ldm
    if "f32" in want:
        outs["f32"] = torch.full((M, N), 7.0, device=dev) if f32_init is None else f32_init.clone()
        e.out_f32, e.ld_f32 = outs["f32"].data_ptr(), N
    e.bias, e.act = _lib.ptr(bias), act
    if gate is not None:
        e.gate, e.ldg = gate.data_ptr(), gate.stride(0)
    e.accumulate_f32 = accumulate
    _lib.check(lib.neraf_gemm_bf16(M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), C.byref(e),
                                   _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    return outs


def _bf16_padded(x, ld):
    out = torch.zeros(x.shape[0], ld, dtype=torch.bfloat16, device=x.device)
    out[:, :x.shape[1]] = x.to(torch.bfloat16)
    return out


@pytest.mark.parametrize("M,N,K", [
    (128, 256, 64), (128, 128, 128), (24, 513, 512), (256, 5096, 163), (300, 2048, 5096), (2048, 1024, 2048),
    (513, 512, 2048), (5096, 163, 256), (128, 64, 64), (1, 8, 8), (2048, 513, 512), (4096, 5096, 168)])
def test_gemm_bf16_tcgen05_matches_torch(M, N, K):
    dev = cuda()
    g = torch.Generator(device="cpu").manual_seed(M + 7 * N + 13 * K)
    A32 = torch.randn(M, K, generator=g).to(dev)
    B32 = torch.randn(N, K, generator=g).to(dev) / np.sqrt(K)
    ldk = (K + 7) // 8 * 8
    A, B = _bf16_padded(A32, ldk), _bf16_padded(B32, ldk)
    if ldk > K:                                  # poison the padding: TMA must zero-fill, not read it
        A[:, K:] = float("nan")
        B[:, K:] = float("nan")
    ref = A[:, :K].double() @ B[:, :K].double().t()
    outs = _gemm_bf16(A, B, M, N, K)
    scale = float(ref.abs().max())
    assert rel_fro(outs["f32"], ref) < 1e-5, "fp32 accumulate of exact bf16 products"
    assert float((outs["rm"][:, :N].double() - ref).abs().max()) < 8e-3 * scale
    assert float((outs["t"][:, :M].double() - ref.t()).abs().max()) < 8e-3 * scale
    if (N + 7) // 8 * 8 > N:
        assert torch.all(outs["rm"][:, N:] == 7.0), "row-major store wrote past N"
    if (M + 7) // 8 * 8 > M:
        assert torch.all(outs["t"][:, M:] == 7.0), "transposed store wrote past M"


@pytest.mark.parametrize("tile", [1064, 1128, 1256, 2064, 2128, 2256])
@pytest.mark.parametrize("M,N,K", [(2048, 2048, 1024), (300, 520, 200), (5096, 163, 2048), (129, 257, 65)])
def test_gemm_bf16_every_tile_shape(tile, M, N, K):
    """Both cta_group variants and all tile widths give the same (correct) result, incl. ragged edges."""
    dev = cuda()
    lib = _lib.lib()
    g = torch.Generator(device="cpu").manual_seed(tile + M)
    A32 = torch.randn(M, K, generator=g).to(dev)
    B32 = torch.randn(N, K, generator=g).to(dev) / np.sqrt(K)
    ldk = (K + 7) // 8 * 8
    A, B = _bf16_padded(A32, ldk), _bf16_padded(B32, ldk)
    ref = A[:, :K].double() @ B[:, :K].double().t()
    _lib.check(lib.neraf_gemm_bf16_set_tile(tile))
    try:
        outs = _gemm_bf16(A, B, M, N, K)
    finally:
        _lib.check(lib.neraf_gemm_bf16_set_tile(0))
    assert rel_fro(outs["f32"], ref) < 1e-5
    assert rel_fro(outs["rm"][:, :N], ref) < 4e-3
    assert rel_fro(outs["t"][:, :M], ref.t()) < 4e-3


@pytest.mark.parametrize("act", [0, 1, 2])
def test_gemm_bf16_epilogue(act):
    dev = cuda()
    M, N, K = 333, 520, 200
    g = torch.Generator(device="cpu").manual_seed(act)
    A = torch.randn(M, K, generator=g).to(dev).to(torch.bfloat16)
    B = (torch.randn(N, K, generator=g) / np.sqrt(K)).to(dev).to(torch.bfloat16)
    bias = torch.randn(N, generator=g).to(dev)
    gate = torch.randn(M, N, generator=g).to(dev).to(torch.bfloat16)
    ref = _act(A.double() @ B.double().t() + bias.double(), act) * torch.where(gate.double() > 0, 1.0, 0.1)
    outs = _gemm_bf16(A, B, M, N, K, bias=bias, act=act, gate=gate)
    assert rel_fro(outs["f32"], ref) < 2e-5
    assert rel_fro(outs["rm"][:, :N], ref) < 4e-3
    assert rel_fro(outs["t"][:, :M], ref.t()) < 4e-3
    init = torch.randn(M, N, generator=g).to(dev)
    outs = _gemm_bf16(A, B, M, N, K, want=("f32",), f32_init=init, accumulate=1)
    assert rel_fro(outs["f32"], A.double() @ B.double().t() + init.double()) < 2e-5


def test_gemm_bf16_rejects_bad_strides():
    dev = cuda()
    A = torch.zeros(16, 12, dtype=torch.bfloat16, device=dev)
    with pytest.raises(_lib.NerafError):
        _gemm_bf16(A, A, 16, 16, 12)             # row stride 12 is not a multiple of 8


def test_convert_bf16_and_transpose():
    dev = cuda()
    lib = _lib.lib()
    x = torch.randn(77, 1187, device=dev)
    out = torch.zeros(77, 1192, dtype=torch.bfloat16, device=dev)
    out_t = torch.zeros(1187, 80, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.neraf_convert_bf16(x.data_ptr(), 77, 1187, 1187, out.data_ptr(), 1192, out_t.data_ptr(), 80,
                                      _lib.stream_ptr(dev)))
    assert torch.equal(out[:, :1187], x.to(torch.bfloat16))
    assert torch.equal(out_t[:, :77], x.to(torch.bfloat16).t())
    assert torch.all(out[:, 1187:] == 0) and torch.all(out_t[:, 77:] == 0)


@pytest.mark.parametrize("shape", [syn.RAF, syn.SOUNDSPACES])
@pytest.mark.parametrize("order", [0, 1])
def test_encode_queries_matches_oracle(shape, order):
    from neraf_b200.field import encode_queries
    dev = cuda()
    batch = syn.make_batch(shape, 1000, seed=11, outside_frac=0.05)
    aabb = syn.default_aabb()
    ref = oenc.encode_queries(batch, aabb, shape.T)
    if order == 1:
        ref = torch.cat([ref[:, 21:84], ref[:, 84:147], ref[:, :21], ref[:, 147:]], dim=-1)
    got = encode_queries(batch["time_query"], batch["mic_pose"], batch["source_pose"], batch["rot"], aabb.to(dev),
                         shape.T, order).cpu()
    assert got.shape == (1000, 163)
    assert torch.equal(got[:, 147:], ref[:, 147:]), "spherical harmonics must be bit-exact (fp16-rounded)"
    assert float((got - ref).abs().max()) < 5e-7          # sin() of identical fp64/fp32 arguments
    rows_out = ~(((oenc.normalize_positions(batch["mic_pose"], aabb) > 0) &
                  (oenc.normalize_positions(batch["mic_pose"], aabb) < 1)).all(-1))
    assert rows_out.any()
    mic0 = 21 if order == 0 else 0
    assert torch.all(got[rows_out][:, mic0 + 60:mic0 + 63] == 0)


def test_encode_empty_batch():
    from neraf_b200.field import encode_queries
    dev = cuda()
    z = torch.zeros(0, 3, dtype=torch.float64)
    out = encode_queries(torch.zeros(0, dtype=torch.int64), z, z, z, syn.default_aabb().to(dev), 60)
    assert out.shape == (0, 163)


@pytest.mark.parametrize("criterion", ["SC+SLMSE", "SC+SLL1", "MSE"])
@pytest.mark.parametrize("n", [(1, 1, 5), (24, 2, 257), (2048, 1, 513)])
def test_spectral_loss_matches_oracle(criterion, n):
    from neraf_b200.loss import spectral_loss
    dev = cuda()
    g = torch.Generator().manual_seed(sum(n))
    x = (torch.randn(*n, generator=g) * 2 - 1).to(dev).requires_grad_(True)
    y = (torch.randn(*n, generator=g) * 2 - 2).to(dev)
    sc, mag = spectral_loss(x, y, criterion, 0.1 * 1e-3, 1e-3)
    ref = oloss.loss_dict(x.detach().cpu(), y.cpu(), criterion, 1e-3)
    if criterion == "MSE":
        assert float(sc) == 0.0
        assert abs(float(mag) - float(ref["audio_mse"])) < 1e-5 * float(ref["audio_mse"])
    else:
        assert abs(float(sc) - float(ref["audio_sc_loss"])) < 1e-5 * float(ref["audio_sc_loss"])
        assert abs(float(mag) - float(ref["audio_mag_loss"])) < 1e-5 * float(ref["audio_mag_loss"])
    (3.0 * sc + 0.5 * mag).backward()
    gref = oloss.loss_grad(x.detach().cpu(), y.cpu(), criterion, 1e-3, g_sc=3.0, g_mag=0.5)
    assert rel_fro(x.grad, gref) < 1e-5
    assert rel_max(x.grad, gref) < 1e-5


def test_spectral_loss_golden_and_module(golden_dir):
    import os
    from neraf_b200.loss import STFTLoss
    dev = cuda()
    gold = np.load(os.path.join(golden_dir, "loss.npz"))
    x = torch.from_numpy(gold["pred"]).to(dev).requires_grad_(True)
    y = torch.from_numpy(gold["gt"]).to(dev)
    for lt in ("mse", "l1"):
        x.grad = None
        out = STFTLoss(loss_type=lt)(x, y)
        assert set(out) == {"audio_sc_loss", "audio_mag_loss"}
        assert abs(float(out["audio_sc_loss"]) - float(gold[f"sc_{lt}"])) < 1e-5 * float(gold[f"sc_{lt}"])
        assert abs(float(out["audio_mag_loss"]) - float(gold[f"mag_{lt}"])) < 1e-5 * float(gold[f"mag_{lt}"])
        (out["audio_sc_loss"] * 1e-4 + out["audio_mag_loss"] * 1e-3).backward()
        assert rel_fro(x.grad, gold[f"grad_{lt}"]) < 1e-5


def test_spectral_loss_full_size_property():
    """At BASELINE size: loss(x, x) == 0 and scale property sc(x, y) unchanged by duplicating the batch."""
    from neraf_b200.loss import spectral_loss
    dev = cuda()
    batch = syn.make_batch(syn.RAF, 2048, seed=3)
    y = batch["data"].to(dev)
    x = (y + 0.3 * torch.randn_like(y)).requires_grad_(True)
    sc0, mag0 = spectral_loss(y.clone(), y, "SC+SLMSE")
    assert float(sc0) == 0.0 and float(mag0) == 0.0
    sc1, mag1 = spectral_loss(x, y, "SC+SLMSE")
    sc2, mag2 = spectral_loss(torch.cat([x, x]), torch.cat([y, y]), "SC+SLMSE")
    assert abs(float(sc1) - float(sc2)) < 1e-6 * float(sc1)
    assert abs(float(mag1) - float(mag2)) < 1e-6 * float(mag1)


def test_loss_rejects_cpu_tensors():
    from neraf_b200.loss import spectral_loss
    with pytest.raises(_lib.NerafError):
        spectral_loss(torch.zeros(4), torch.zeros(4))


# ------------------------------------------------------------------------------------------------
# job-list kernel (neraf_gemm_bf16_jobs): one persistent launch, dependent GEMMs, MN-major operands
# ------------------------------------------------------------------------------------------------
def _run_jobs(jobs, dev):
    lib = _lib.lib()
    arr = (_lib.GemmJob * len(jobs))(*jobs)
    counters = torch.empty(4096, dtype=torch.int32, device=dev)
    _lib.check(lib.neraf_gemm_bf16_jobs(arr, len(jobs), counters.data_ptr(), counters.numel() * 4, _lib.stream_ptr(dev)))
    torch.cuda.synchronize()


def _job(M, N, K, A, B, bn, a_mn=0, b_mn=0, wait_job=-1, wait_all=0):
    j = _lib.GemmJob()
    j.M, j.N, j.K, j.A, j.lda, j.B, j.ldb = M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0)
    j.a_mn, j.b_mn, j.bn, j.wait_job, j.wait_all = a_mn, b_mn, bn, wait_job, wait_all
    return j


@pytest.mark.parametrize("bn", [64, 128, 256])
@pytest.mark.parametrize("M,N,K", [(2048, 1024, 512), (300, 513, 200), (256, 64, 64), (1000, 5096, 168)])
def test_gemm_jobs_k_major_outputs(bn, M, N, K):
    dev = cuda()
    g = torch.Generator().manual_seed(bn + M)
    ldk = (K + 7) // 8 * 8
    A = _bf16_padded(torch.randn(M, K, generator=g).to(dev), ldk)
    B = _bf16_padded((torch.randn(N, K, generator=g) / np.sqrt(K)).to(dev), ldk)
    bias = torch.randn(N, generator=g).to(dev)
    gate = torch.randn(M, (N + 7) // 8 * 8, generator=g).to(dev).to(torch.bfloat16)
    ref = torch.nn.functional.leaky_relu(A[:, :K].double() @ B[:, :K].double().t() + bias.double(), 0.1)
    ref = ref * torch.where(gate[:, :N].double() > 0, 1.0, 0.1)
    # job 0: bf16 output (TMA store) + column sums; job 1: same GEMM, fp32 output with an unaligned row stride
    n8 = (N + 7) // 8 * 8
    out_b = torch.full((M, n8 + 8), 7.0, dtype=torch.bfloat16, device=dev)
    colsum = torch.zeros(N, device=dev)
    out_f = torch.full((M, N + 1), 7.0, device=dev)
    j0 = _job(M, N, K, A, B, bn)
    j0.epi.bias, j0.epi.act, j0.epi.gate, j0.epi.ldg = bias.data_ptr(), 1, gate.data_ptr(), gate.stride(0)
    j0.epi.out_bf16, j0.epi.ld_bf16, j0.colsum = out_b.data_ptr(), out_b.stride(0), colsum.data_ptr()
    j1 = _job(M, N, K, A, B, bn)
    j1.epi.bias, j1.epi.act, j1.epi.gate, j1.epi.ldg = bias.data_ptr(), 1, gate.data_ptr(), gate.stride(0)
    j1.epi.out_f32, j1.epi.ld_f32 = out_f.data_ptr(), out_f.stride(0)
    # job 2: fp32 output through the aligned (TMA) path
    out_f2 = torch.full((M, (N + 3) // 4 * 4), 7.0, device=dev)
    j2 = _job(M, N, K, A, B, bn)
    j2.epi.out_f32, j2.epi.ld_f32 = out_f2.data_ptr(), out_f2.stride(0)
    _run_jobs([j0, j1, j2], dev)
    assert rel_fro(out_f[:, :N], ref) < 2e-5
    assert torch.all(out_f[:, N] == 7.0)
    assert rel_fro(out_b[:, :N], ref) < 4e-3
    # bf16 rows go out through TMA stores, which clip with 16-byte granularity: pad columns sharing the last valid
    # column's 16-byte group may receive zeros, nothing beyond round_up(N, 8) is touched.  fp32 rows are exact.
    assert torch.all(out_b[:, n8:] == 7.0), "TMA store must clip at round_up(N, 8)"
    pad = out_b[:, N:n8]
    assert torch.all((pad == 7.0) | (pad == 0.0))
    assert torch.all(out_f2[:, N:] == 7.0)
    assert rel_fro(colsum, ref.sum(0)) < 1e-4
    assert rel_fro(out_f2[:, :N], A[:, :K].double() @ B[:, :K].double().t()) < 2e-5


@pytest.mark.parametrize("bn", [128, 256])
@pytest.mark.parametrize("n_out,k_in,batch", [(512, 1024, 2048), (513, 512, 300), (5096, 163, 1000), (200, 130, 64)])
def test_gemm_jobs_mn_major_weight_gradient(bn, n_out, k_in, batch):
    """dW = dZ^T X with both operands read MN-major straight from their row-major (batch, features) storage."""
    dev = cuda()
    g = torch.Generator().manual_seed(n_out + batch)
    dz = _bf16_padded(torch.randn(batch, n_out, generator=g).to(dev), (n_out + 7) // 8 * 8)
    x = _bf16_padded(torch.randn(batch, k_in, generator=g).to(dev), (k_in + 7) // 8 * 8)
    ref = dz[:, :n_out].double().t() @ x[:, :k_in].double()
    for ld in (k_in, (k_in + 3) // 4 * 4 + 4):          # unaligned (manual store) and 16-byte aligned (TMA store) strides
        out = torch.full((n_out, ld), 7.0, device=dev)
        j = _job(n_out, k_in, batch, dz, x, bn, a_mn=1, b_mn=1)
        j.epi.out_f32, j.epi.ld_f32 = out.data_ptr(), ld
        _run_jobs([j], dev)
        assert rel_fro(out[:, :k_in], ref) < 2e-5, ld
        if ld > k_in:
            assert torch.all(out[:, k_in:] == 7.0)


@pytest.mark.parametrize("M,N,K,bn", [(2048, 1024, 256, 128), (300, 513, 200, 64), (1000, 5096, 168, 256)])
def test_gemm_jobs_sign_mask_gate(M, N, K, bn):
    """Forward job stores the transposed sign bit mask; a second job gated by the mask == gated by the activations."""
    dev = cuda()
    g = torch.Generator().manual_seed(M + N)
    ldk = (K + 7) // 8 * 8
    n8 = (N + 7) // 8 * 8
    A = _bf16_padded(torch.randn(M, K, generator=g).to(dev), ldk)
    B = _bf16_padded((torch.randn(N, K, generator=g) / np.sqrt(K)).to(dev), ldk)
    bias = torch.randn(N, generator=g).to(dev) * 0.1
    ld_mask = (M + 63) // 64 * 64
    mask = torch.zeros((N + 31) // 32, ld_mask, dtype=torch.int32, device=dev)
    x = torch.zeros(M, n8, dtype=torch.bfloat16, device=dev)
    j0 = _job(M, N, K, A, B, bn)
    j0.epi.bias, j0.epi.act = bias.data_ptr(), 1
    j0.epi.out_bf16, j0.epi.ld_bf16 = x.data_ptr(), n8
    j0.epi.mask_out, j0.epi.ld_mask = mask.data_ptr(), ld_mask
    out_m = torch.zeros(M, N + 1, device=dev)
    out_g = torch.zeros(M, N + 1, device=dev)
    j1 = _job(M, N, K, A, B, bn, wait_job=0)
    j1.epi.gate_mask, j1.epi.ld_mask = mask.data_ptr(), ld_mask
    j1.epi.out_f32, j1.epi.ld_f32 = out_m.data_ptr(), N + 1
    j2 = _job(M, N, K, A, B, bn, wait_job=0)
    j2.epi.gate, j2.epi.ldg = x.data_ptr(), n8
    j2.epi.out_f32, j2.epi.ld_f32 = out_g.data_ptr(), N + 1
    _run_jobs([j0, j1, j2], dev)
    z = A[:, :K].double() @ B[:, :K].double().t() + bias.double()
    cols = torch.arange(32, device=dev)
    shifts = (cols & 1) * 16 + (cols >> 1)                  # bit i <- column 2i, bit 16 + i <- column 2i + 1
    bits = (mask[:, :M].t().unsqueeze(-1) >> shifts) & 1                               # (M, words, 32), 1 = negative
    pos = ~bits.reshape(M, -1)[:, :N].bool()
    sure = z.abs() > 1e-4                                   # away from zero the fp32 accumulator has the sign of z
    assert torch.all(pos[sure] == (z[sure] > 0))
    # the bf16 activation has the sign of its fp32 source, so both gates agree wherever x did not round to zero
    agree = (out_m[:, :N] == out_g[:, :N]) | (x[:, :N] == 0)
    assert torch.all(agree)
    ref = (A[:, :K].double() @ B[:, :K].double().t()) * torch.where(pos, 1.0, 0.1)
    assert rel_fro(out_m[:, :N], ref) < 2e-5


def test_gemm_jobs_dependency_chain_matches_sequential():
    """Three chained layers + a weight gradient that needs ALL row blocks of the middle layer, one launch."""
    dev = cuda()
    g = torch.Generator().manual_seed(5)
    Bsz, d0, d1, d2, d3 = 1536, 168, 1024, 512, 264
    x0 = torch.randn(Bsz, d0, generator=g).to(dev).to(torch.bfloat16)
    w1 = (torch.randn(d1, d0, generator=g) / np.sqrt(d0)).to(dev).to(torch.bfloat16)
    w2 = (torch.randn(d2, d1, generator=g) / np.sqrt(d1)).to(dev).to(torch.bfloat16)
    w3 = (torch.randn(d3, d2, generator=g) / np.sqrt(d2)).to(dev).to(torch.bfloat16)
    x1 = torch.zeros(Bsz, d1, dtype=torch.bfloat16, device=dev)
    x2 = torch.zeros(Bsz, d2, dtype=torch.bfloat16, device=dev)
    y = torch.zeros(Bsz, d3, device=dev)
    dw = torch.zeros(d2, d1, device=dev)
    j0 = _job(Bsz, d1, d0, x0, w1, 128); j0.epi.act = 1; j0.epi.out_bf16, j0.epi.ld_bf16 = x1.data_ptr(), d1
    j1 = _job(Bsz, d2, d1, x1, w2, 64, wait_job=0); j1.epi.act = 1; j1.epi.out_bf16, j1.epi.ld_bf16 = x2.data_ptr(), d2
    j2 = _job(Bsz, d3, d2, x2, w3, 64, wait_job=1); j2.epi.act = 2; j2.epi.out_f32, j2.epi.ld_f32 = y.data_ptr(), d3
    j3 = _job(d2, d1, Bsz, x2, x1, 128, a_mn=1, b_mn=1, wait_job=1, wait_all=1); j3.epi.out_f32, j3.epi.ld_f32 = dw.data_ptr(), d1
    for _ in range(3):                               # repeat: counters must be reset by every launch
        _run_jobs([j0, j1, j2, j3], dev)
    lr = torch.nn.functional.leaky_relu
    r1 = lr(x0.float() @ w1.float().t(), 0.1).to(torch.bfloat16)
    r2 = lr(r1.float() @ w2.float().t(), 0.1).to(torch.bfloat16)
    ry = 10 * torch.tanh(r2.float() @ w3.float().t())
    assert rel_fro(x1, r1) < 3e-3 and rel_fro(x2, r2) < 4e-3
    assert rel_fro(y, ry) < 1e-2
    assert rel_fro(dw, x2.double().t() @ x1.double()) < 2e-5


def test_gemm_jobs_interleaved_pair():
    """merge_next: the tiles of two independent jobs alternate in the tile list; results are those of the plain list."""
    dev = cuda()
    g = torch.Generator().manual_seed(11)
    M, K = 1024, 256
    A = torch.randn(M, K, generator=g).to(dev).to(torch.bfloat16)
    B1 = (torch.randn(1000, K, generator=g) / 16).to(dev).to(torch.bfloat16)
    B2 = (torch.randn(328, K, generator=g) / 16).to(dev).to(torch.bfloat16)
    x0 = torch.zeros(M, 512, dtype=torch.bfloat16, device=dev)
    W0 = (torch.randn(512, K, generator=g) / 16).to(dev).to(torch.bfloat16)
    o1, o2 = torch.zeros(M, 1000, device=dev), torch.zeros(M, 328, device=dev)
    j0 = _job(M, 512, K, A, W0, 128); j0.epi.out_bf16, j0.epi.ld_bf16 = x0.data_ptr(), 512
    j1 = _job(M, 1000, K, A, B1, 128, wait_job=0); j1.epi.out_f32, j1.epi.ld_f32 = o1.data_ptr(), 1000
    j1.merge_next = 1
    j2 = _job(M, 328, K, A, B2, 64, wait_job=0, wait_all=1); j2.epi.out_f32, j2.epi.ld_f32 = o2.data_ptr(), 328
    _run_jobs([j0, j1, j2], dev)
    assert rel_fro(o1, A.double() @ B1.double().t()) < 2e-5
    assert rel_fro(o2, A.double() @ B2.double().t()) < 2e-5
    assert rel_fro(x0, A.double() @ W0.double().t()) < 4e-3
    j2.wait_job = 1                                   # the second job of a pair must not wait for the first
    with pytest.raises(_lib.NerafError):
        _run_jobs([j0, j1, j2], dev)


def test_gemm_jobs_validation():
    dev = cuda()
    A = torch.zeros(64, 64, dtype=torch.bfloat16, device=dev)
    out = torch.zeros(64, 64, device=dev)
    j = _job(64, 64, 64, A, A, 64, a_mn=1, b_mn=1)       # MN-major B needs bn >= 128
    j.epi.out_f32, j.epi.ld_f32 = out.data_ptr(), 64
    with pytest.raises(_lib.NerafError):
        _run_jobs([j], dev)
    j = _job(64, 64, 64, A, A, 64, wait_job=0)          # depends on itself
    j.epi.out_f32, j.epi.ld_f32 = out.data_ptr(), 64
    with pytest.raises(_lib.NerafError):
        _run_jobs([j], dev)
