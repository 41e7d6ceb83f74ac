"""Grid-feature producer, CPU tests.

* the oracle (oracle/gridnet.py) against the golden vectors produced by the reference's real ResNet3D_helper;
* the per-element code of the CUDA kernels (csrc/gridnet_core.h, built for the host) against torch's own operators,
  on non-cubic extents so that an axis mix-up cannot hide;
* neraf_b200/gridnet.py's assembly of the whole network, run on those host-built operators, against the oracle.

Tolerances.  Evaluation mode (running statistics: every layer is well conditioned) -- fp32: feature 1e-6, every
gradient tensor 5e-3 with a median of 1e-5 (ONE ReLU gate that flips between fp32 and the float64 oracle in a
512-voxel unit moves every tensor behind it by ~1e-3: seen on the B200, profiles/r01f_gridnet_gpu_tests.txt); bf16: feature
1e-2 (north_star's bf16 bound), gradients 5e-2 median.  Training mode normalises with statistics of the batch of ONE
grid: the gradient through batch-norm + global average pooling is the small remainder of a cancellation, so torch's
own fp32 gradient is already 1.5e-2 away from its float64 value on this problem and its bf16-autocast gradient is
noise (measured in DESIGN.md section 9).  There the bound is relative to torch's own error at the same precision,
computed in the test.
"""
import json
import os
import statistics

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from neraf_b200 import _lib, synthetic as syn
from neraf_b200.gridnet import ResNet3D_helper, Window3d
from oracle import gridnet as og
from tests.util import rel_fro

N, GRID_STEP = 64, 1 / 64


@pytest.fixture(scope="module")
def host_ops(built):
    from tests.gridnet_host import HostOps
    return HostOps()


@pytest.fixture(scope="module")
def problem():
    sd = syn.make_gridnet_state_dict("resnet50")
    x = syn.make_grid(N)
    dout = torch.randn(1, 1024, 1, 1, 1, generator=torch.Generator().manual_seed(5))
    return sd, x, dout


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "gridnet_resnet50.npz"))


def _projection(t, name):
    seed = sum((i + 1) * ord(ch) for i, ch in enumerate(name)) % (2 ** 31)
    g = torch.Generator().manual_seed(seed)
    return float((t.double() * torch.randn(t.shape, generator=g).double()).sum())


# ------------------------------------------------------------------------------------------------ oracle vs reference
def test_oracle_matches_the_reference_golden(problem, golden):
    # fp32 on the host: torch picks its convolution / reduction kernels per ISA (the golden file was written on an AVX2
    # host; on an AVX512 host the same 50 layers land 1.1e-6 away), so the gates are fp32 round-off across hosts, not
    # bit equality -- the float64 comparisons below (and the GPU tests) are the tight ones
    sd, x, dout = problem
    assert np.array_equal(golden["dout"], dout.reshape(-1).numpy())
    stats = {}
    with torch.no_grad():
        f_train = og.forward(sd, x, GRID_STEP, 1024, True, new_stats=stats)
    assert rel_fro(f_train.reshape(-1), golden["feature_train"]) < 5e-6
    for k in golden.files:
        if k.startswith("stats/"):
            assert rel_fro(stats[k[6:]], golden[k]) < 5e-6, k
    f_eval, grads, _ = og.forward_backward(sd, x, dout, GRID_STEP, training=False, dtype=torch.float32)
    assert rel_fro(f_eval.reshape(-1), golden["feature_eval"]) < 5e-6
    norms = json.loads(str(golden["grad_eval_norms"]))
    projs = json.loads(str(golden["grad_eval_projections"]))
    assert set(norms) == set(grads)
    for k, g in grads.items():
        assert abs(float(g.double().norm()) - norms[k]) <= 1e-4 * norms[k], k
        assert abs(_projection(g, k) - projs[k]) <= 1e-4 * norms[k], k
    for k in golden.files:
        if k.startswith("grad_eval/"):
            assert rel_fro(grads[k[10:]], golden[k]) < 5e-5, k


def test_state_dict_is_the_reference_checkpoint_contract(golden):
    want = json.loads(str(golden["state_keys"]))
    net = ResNet3D_helper(in_channels=7, backbone="resnet50", pretrained=False, grid_step=GRID_STEP, N_features=1024)
    got = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert got == want
    assert list(got) == list(want), "same order: optimizers index parameters by position"
    assert net.backbone_net.avgpool_size == og.avgpool_window(GRID_STEP, 1024) == 4
    for step, n_feat, size in ((1 / 128, 1024, 8), (1 / 128, 2048, 4), (1 / 256, 1024, 16), (None, 1024, 8)):
        assert ResNet3D_helper(7, "resnet18", False, step, n_feat).backbone_net.avgpool_size == size


# ------------------------------------------------------------------------------------------------ operators on the host
def _act(t5):                      # (1, C, D, H, W) -> channels-last matrix (V, C)
    c = t5.shape[1]
    return t5[0].permute(1, 2, 3, 0).reshape(-1, c).contiguous()


def _unact(m, dims):               # (V, C) -> (1, C, D, H, W)
    return m.reshape(*dims, m.shape[1]).permute(3, 0, 1, 2)[None].contiguous()


@pytest.mark.parametrize("dims,c_in,c_out,k,stride,pad,from_grid", [
    ((9, 6, 7), 7, 16, 5, 2, 2, True),          # the stem, read from the channels-first grid
    ((6, 5, 7), 8, 24, 3, 1, 1, False),
    ((7, 6, 5), 16, 8, 3, 2, 1, False),
    ((6, 4, 5), 8, 16, 1, 2, 0, False),         # the shortcut convolution
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_convolution_as_gather_and_gemm(host_ops, dims, c_in, c_out, k, stride, pad, from_grid, dtype):
    g = torch.Generator().manual_seed(k * 100 + stride)
    x = torch.randn(1, c_in, *dims, generator=g)
    wt = torch.randn(c_out, c_in, k, k, k, generator=g) / np.sqrt(c_in * k ** 3)
    if dtype == torch.bfloat16 and not from_grid:
        x = x.bfloat16().float()
    xr = x.double().requires_grad_(True)
    wr = (wt.bfloat16().double() if dtype == torch.bfloat16 else wt.double()).requires_grad_(True)
    y_ref = F.conv3d(xr, wr, stride=stride, padding=pad)
    dy = torch.randn(y_ref.shape, generator=g)
    if dtype == torch.bfloat16:
        dy = dy.bfloat16().float()
    y_ref.backward(dy.double())

    w = Window3d(dims[0], dims[1], dims[2], c_in, k, stride, pad)
    od = w.out_dims
    assert tuple(y_ref.shape[2:]) == od
    v_out, kc = od[0] * od[1] * od[2], k ** 3 * c_in
    ld = (kc + 7) // 8 * 8
    col = torch.full((v_out, ld), 9.0, dtype=dtype)
    if from_grid:
        host_ops.im2col(w, x.contiguous(), 1, dims[0] * dims[1] * dims[2], col)
    else:
        host_ops.im2col(w, _act(x).to(dtype), c_in, 1, col)
    assert torch.all(col[:, kc:] == 0), "pad columns are zero-filled"
    wmat = torch.full((c_out, ld), 9.0, dtype=dtype)
    host_ops.pack_weight(wt, wmat)
    assert torch.all(wmat[:, kc:] == 0)
    y = torch.empty(v_out, c_out, dtype=torch.float32)
    host_ops.gemm_nt(col, wmat, v_out, c_out, kc, y)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_fro(_unact(y, od), y_ref) < tol
    # weight gradient: contraction over the voxels, then back to the parameter's layout
    dym = _act(dy).to(dtype)
    dw_mat = torch.zeros(c_out, ld)
    host_ops.gemm_tn(dym, col, c_out, kc, v_out, dw_mat)
    dw = torch.empty_like(wt)
    host_ops.unpack_wgrad(dw_mat, dw)
    assert rel_fro(dw, wr.grad) < tol
    # data gradient: GEMM + gather
    dcol = torch.zeros(v_out, ld, dtype=dtype)
    host_ops.gemm_nn(dym, wmat, v_out, kc, c_out, dcol)
    dx = torch.full((dims[0] * dims[1] * dims[2], c_in), 9.0, dtype=dtype)
    host_ops.col2im(w, dcol, dx)
    assert rel_fro(_unact(dx.float(), dims), xr.grad) < (1e-5 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_maxpool_with_ties_routes_gradients_like_torch(host_ops, dtype):
    g = torch.Generator().manual_seed(3)
    dims, c = (7, 6, 9), 8
    x = torch.relu(torch.randn(1, c, *dims, generator=g))            # ~half the entries tie at zero, as after a ReLU
    x = torch.round(x * 4) / 4                                       # and positive values tie too
    xr = x.clone().requires_grad_(True)
    y_ref = F.max_pool3d(xr, 3, 2, 1)
    dy = torch.round(torch.randn(y_ref.shape, generator=g) * 8) / 8
    dy2 = torch.round(torch.randn(y_ref.shape, generator=g) * 8) / 8
    y_ref.backward(dy + dy2)
    w = Window3d(dims[0], dims[1], dims[2], c, 3, 2, 1)
    od = w.out_dims
    v_out = od[0] * od[1] * od[2]
    y = torch.empty(v_out, c, dtype=dtype)
    arg = torch.empty(v_out, c, dtype=torch.int32)
    host_ops.maxpool(w, _act(x).to(dtype), y, arg)
    assert torch.equal(_unact(y.float(), od), y_ref.detach())
    dx = torch.empty(dims[0] * dims[1] * dims[2], c, dtype=dtype)
    host_ops.maxpool_backward(w, _act(dy).to(dtype), _act(dy2).to(dtype), arg, dx)
    assert torch.equal(_unact(dx.float(), dims), xr.grad)
    host_ops.maxpool_backward(w, _act(dy).to(dtype), None, arg, dx)
    xr.grad = None
    F.max_pool3d(xr, 3, 2, 1).backward(dy)
    assert torch.equal(_unact(dx.float(), dims), xr.grad)


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("relu,with_res", [(True, True), (True, False), (False, False)])
def test_batchnorm_unit_forward_and_backward(host_ops, training, relu, with_res):
    g = torch.Generator().manual_seed(11)
    V, c = 333, 24
    x = torch.randn(V, c, generator=g) * 2 + 0.5
    res = torch.randn(V, c, generator=g) if with_res else None
    gamma, beta = 1 + 0.2 * torch.randn(c, generator=g), 0.2 * torch.randn(c, generator=g)
    rm, rv = 0.1 * torch.randn(c, generator=g), 0.5 + torch.rand(c, generator=g)
    dy, dy2 = torch.randn(V, c, generator=g), torch.randn(V, c, generator=g)
    # torch, float64
    xr, gr, br = x.double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm_ref, rv_ref = rm.double().clone(), rv.double().clone()
    y_ref = F.batch_norm(xr.t()[None], rm_ref, rv_ref, gr, br, training, 0.1, 1e-5)[0].t()
    if with_res:
        resr = res.double().requires_grad_(True)
        y_ref = y_ref + resr
    if relu:
        y_ref = F.relu(y_ref)
    y_ref.backward((dy + dy2).double())
    # host operators
    sums = torch.empty(2, c, dtype=torch.float64)
    mean, invstd = torch.empty(c), torch.empty(c)
    rm_h, rv_h = rm.clone(), rv.clone()
    if training:
        host_ops.bn_stats(x, sums)
        assert rel_fro(sums[0], x.double().sum(0)) < 1e-6 and rel_fro(sums[1], (x.double() ** 2).sum(0)) < 1e-6
    host_ops.bn_finalize(sums if training else None, V, c, 1e-5, 0.1 if training else 0.0, training, rm_h, rv_h, mean, invstd)
    assert rel_fro(rm_h, rm_ref) < 1e-6 and rel_fro(rv_h, rv_ref) < 1e-6
    y = torch.empty(V, c)
    host_ops.bn_apply(x, mean, invstd, gamma, beta, res, relu, y)
    assert rel_fro(y, y_ref) < 1e-6
    gbuf, dx = torch.empty(V, c), torch.empty(V, c)
    dgamma, dbeta = torch.empty(c), torch.empty(c)
    host_ops.bn_backward_reduce(dy, dy2, y if relu else None, x, mean, invstd, gbuf, sums)
    host_ops.bn_backward_apply(gbuf, x, mean, invstd, gamma, sums, training, dx, dgamma, dbeta)
    assert rel_fro(dx, xr.grad) < 1e-5
    assert rel_fro(dgamma, gr.grad) < 1e-5 and rel_fro(dbeta, br.grad) < 1e-5
    if with_res:
        assert rel_fro(gbuf, resr.grad) < 1e-6          # the shortcut receives exactly g
    # in-place form (dx aliases g), as the assembly uses it for units without a shortcut
    host_ops.bn_backward_apply(gbuf, x, mean, invstd, gamma, sums, training, gbuf, None, None)
    assert torch.equal(gbuf, dx)


def test_global_average_pool_and_its_gradient(host_ops):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(64, 40, generator=g).bfloat16()
    sums, feat = torch.empty(2, 40, dtype=torch.float64), torch.empty(40)
    host_ops.bn_stats(x, sums)
    host_ops.bn_finalize(sums, 64, 40, 0.0, 0.0, True, None, None, feat, None)
    assert rel_fro(feat, x.double().mean(0)) < 1e-6
    d = torch.empty(64, 40, dtype=torch.bfloat16)
    v = torch.randn(40, generator=g)
    host_ops.broadcast_rows(v, 1 / 64, d)
    assert torch.equal(d, (v / 64).bfloat16()[None].expand(64, 40))


# ------------------------------------------------------------------------------------------------ the whole network
def _run(host_ops, sd, x, dout, precision, training, backbone="resnet50", n_features=1024, grid_step=GRID_STEP):
    net = ResNet3D_helper(in_channels=7, backbone=backbone, grid_step=grid_step, N_features=n_features, precision=precision)
    net.load_state_dict(sd, strict=True)
    net.backbone_net.ops = host_ops
    net.train(training)
    out = net(x)
    out.backward(dout)
    return net, out.detach()


def test_network_eval_mode_fp32_matches_oracle_and_golden(host_ops, problem, golden):
    sd, x, dout = problem
    ref, grads, _ = og.forward_backward(sd, x, dout, GRID_STEP, training=False)
    net, out = _run(host_ops, sd, x, dout, "fp32", False)
    assert out.shape == (1, 1024, 1, 1, 1) and out.dtype == torch.float32
    assert rel_fro(out, ref) < 1e-6
    assert rel_fro(out.reshape(-1), golden["feature_eval"]) < 1e-6
    errs = {k: rel_fro(p.grad, grads[k]) for k, p in net.named_parameters()}
    assert max(errs.values()) < 5e-3, max(errs.items(), key=lambda kv: kv[1])
    assert statistics.median(errs.values()) < 1e-5
    for k in golden.files:
        if k.startswith("grad_eval/"):
            assert rel_fro(dict(net.named_parameters())[k[10:]].grad, golden[k]) < 5e-3, k
    after = net.state_dict()
    assert all(torch.equal(after[k], sd[k]) for k in sd), "evaluation mode leaves the running statistics alone"


def test_network_training_mode_fp32(host_ops, problem, golden):
    sd, x, dout = problem
    ref, grads, stats = og.forward_backward(sd, x, dout, GRID_STEP, training=True)
    _, grads32, _ = og.forward_backward(sd, x, dout, GRID_STEP, training=True, dtype=torch.float32)
    net, out = _run(host_ops, sd, x, dout, "fp32", True)
    assert rel_fro(out, ref) < 2e-5
    assert rel_fro(out.reshape(-1), golden["feature_train"]) < 2e-5
    after = net.state_dict()
    for k, v in stats.items():
        assert rel_fro(after[k], v) < 1e-5, k
    for k in golden.files:
        if k.startswith("stats/"):
            assert rel_fro(after[k[6:]], golden[k]) < 1e-5, k
    assert int(after["backbone_net.layer3.5.bn3.num_batches_tracked"]) == int(golden["num_batches_tracked"]) == 1
    ours = statistics.median(rel_fro(p.grad, grads[k]) for k, p in net.named_parameters())
    torch32 = statistics.median(rel_fro(grads32[k], grads[k]) for k in grads)
    assert torch32 > 1e-3, "the premise of this bound: torch's own fp32 gradient is far from float64 here"
    assert ours < 3 * torch32


@pytest.mark.parametrize("training", [False, True])
def test_network_bf16(host_ops, problem, training):
    sd, x, dout = problem
    ref, grads, _ = og.forward_backward(sd, x, dout, GRID_STEP, training=training)
    net, out = _run(host_ops, sd, x, dout, "bf16", training)
    err = rel_fro(out, ref)
    if not training:
        assert err < 1e-2
        assert statistics.median(rel_fro(p.grad, grads[k]) for k, p in net.named_parameters()) < 5e-2
    else:
        assert err < 5e-2          # torch's own bf16 autocast: 2.3e-2 on this problem
    assert all(p.grad is not None and p.grad.dtype == torch.float32 and torch.isfinite(p.grad).all()
               for p in net.parameters())


def test_basic_block_backbone_and_2048_features(host_ops):
    sd = syn.make_gridnet_state_dict("resnet18", N_features=2048, seed=4)
    x = syn.make_grid(64, seed=1)                # 64 -> 32 -> 16 -> 16 -> 8 -> 4 -> 2, pooling window 2
    dout = torch.randn(1, 512, 1, 1, 1, generator=torch.Generator().manual_seed(6))
    ref, grads, _ = og.forward_backward(sd, x, dout, GRID_STEP, n_features=2048, training=False)
    net, out = _run(host_ops, sd, x, dout, "fp32", False, backbone="resnet18", n_features=2048)
    assert out.shape == ref.shape == (1, 512, 1, 1, 1)
    assert rel_fro(out, ref) < 1e-6
    errs = [rel_fro(p.grad, grads[k]) for k, p in net.named_parameters()]
    assert max(errs) < 5e-3 and statistics.median(errs) < 1e-5


def test_calling_conventions_and_errors(host_ops, problem):
    sd, x, dout = problem
    net = ResNet3D_helper(in_channels=7, backbone="resnet50", grid_step=GRID_STEP, N_features=1024)
    net.load_state_dict(sd)
    with pytest.raises(_lib.NerafError):                     # the product path has no CPU implementation
        net(x)
    net.backbone_net.ops = host_ops
    with pytest.raises(ValueError):
        net(x[0])
    with pytest.raises(ValueError):
        net(x[:, :5])
    with pytest.raises(_lib.NerafError):                     # 32^3 leaves 2^3 voxels, the pooling window is 4^3
        net(syn.make_grid(32))
    with pytest.raises(ValueError):
        ResNet3D_helper(precision="fp16")
    net.eval()
    with torch.no_grad():
        f = net(x)
    assert not f.requires_grad and f.shape == (1, 1024, 1, 1, 1)


def test_model_builds_the_reference_grid_net(host_ops):
    """NeRAF_model.py:185,554-557: the model owns a ResNet3D_helper and feeds it grid[None]; its parameters join the
    "audio_fields" group (:734)."""
    from neraf_b200.model import NeRAFAudioModel, NeRAFAudioModelConfig
    cfg = NeRAFAudioModelConfig(dataset="RAF", grid_step=GRID_STEP, grid_net="resnet50", precision="fp32")
    model = NeRAFAudioModel(cfg, syn.default_aabb(), grid=syn.make_grid(N)[0])
    assert isinstance(model.resnet3d, ResNet3D_helper) and model.resnet3d.backbone_net.precision == "fp32"
    sd = syn.make_gridnet_state_dict("resnet50")
    model.resnet3d.load_state_dict(sd)
    model.resnet3d.backbone_net.ops = host_ops
    model.resnet3d.eval()
    feat = model.grid_feature()
    with torch.no_grad():
        ref = og.forward({k: v.double() if v.dtype.is_floating_point else v for k, v in sd.items()},
                         syn.make_grid(N).double(), GRID_STEP, training=False)
    assert feat.shape == (1024,) and rel_fro(feat, ref.reshape(-1)) < 1e-6
    group = model.get_param_groups()["audio_fields"]
    assert all(any(p is q for q in group) for p in model.resnet3d.parameters())
    with pytest.raises(ValueError):
        NeRAFAudioModel(NeRAFAudioModelConfig(grid_net="vgg"), syn.default_aabb())


def test_16_byte_gathers_equal_the_scalar_forms(host_ops):
    """bf16 activations with a multiple of 8 channels take the 128-bit kernels (gather_can_vec8); a buffer that starts
    2 bytes off a 16-byte boundary forces the scalar kernels.  Same bits either way."""
    g = torch.Generator().manual_seed(9)
    dims, c = (6, 5, 7), 16
    w = Window3d(dims[0], dims[1], dims[2], c, 3, 2, 1)
    od = w.out_dims
    v_in, v_out, kc = dims[0] * dims[1] * dims[2], od[0] * od[1] * od[2], 27 * c
    x = torch.randn(v_in, c, generator=g).bfloat16()
    col_v = torch.empty(v_out, kc, dtype=torch.bfloat16)
    host_ops.im2col(w, x, c, 1, col_v)
    off = torch.empty(v_out * kc + 1, dtype=torch.bfloat16)[1:].view(v_out, kc)
    assert off.data_ptr() % 16 != 0 and col_v.data_ptr() % 16 == 0
    host_ops.im2col(w, x, c, 1, off)
    assert torch.equal(off, col_v)
    dcol = torch.randn(v_out, kc, generator=g).bfloat16()
    dx_v = torch.empty(v_in, c, dtype=torch.bfloat16)
    host_ops.col2im(w, dcol, dx_v)
    dx_s = torch.empty(v_in * c + 1, dtype=torch.bfloat16)[1:].view(v_in, c)
    host_ops.col2im(w, dcol, dx_s)
    assert torch.equal(dx_s, dx_v)


def test_16_byte_batchnorm_passes_equal_the_scalar_forms(host_ops):
    """bn_apply / bn_backward_apply on bf16 rows of whole 16-byte words take the 128-bit kernels; an output buffer 2 bytes
    off a 16-byte boundary forces the scalar kernels.  Same bits."""
    g = torch.Generator().manual_seed(12)
    V, c = 77, 24
    x = (torch.randn(V, c, generator=g) * 2).bfloat16()
    res = torch.randn(V, c, generator=g).bfloat16()
    gam, bet = 1 + 0.2 * torch.randn(c, generator=g), 0.2 * torch.randn(c, generator=g)
    mean, invstd = 0.1 * torch.randn(c, generator=g), 0.5 + torch.rand(c, generator=g)
    sums = torch.randn(2, c, generator=g, dtype=torch.float64)

    def off(rows):                      # (rows, c) bf16 view that starts 2 bytes past a 16-byte boundary, row stride c + 8
        t = torch.empty(rows * (c + 8) + 1, dtype=torch.bfloat16)[1:].view(rows, c + 8)[:, :c]
        assert t.data_ptr() % 16 != 0
        return t

    for r, relu in ((res, 1), (None, 0)):
        y_v, y_s = torch.empty(V, c, dtype=torch.bfloat16), off(V)
        host_ops.bn_apply(x, mean, invstd, gam, bet, r, relu, y_v)
        host_ops.bn_apply(x, mean, invstd, gam, bet, r, relu, y_s)
        assert torch.equal(y_s, y_v)
    gy = torch.randn(V, c, generator=g).bfloat16()
    for training in (True, False):
        d_v = torch.empty(V, c, dtype=torch.bfloat16)
        host_ops.bn_backward_apply(gy, x, mean, invstd, gam, sums, training, d_v, None, None)
        # scalar form: all three matrices must share the row stride, so give them all the offset layout
        gy_s, x_s, d_s = off(V), off(V), off(V)
        gy_s.copy_(gy); x_s.copy_(x)
        host_ops.bn_backward_apply(gy_s, x_s, mean, invstd, gam, sums, training, d_s, None, None)
        assert torch.equal(d_s, d_v)


def test_16_byte_batchnorm_reductions_and_pooling_gradient_equal_the_scalar_forms(host_ops):
    """bn_stats / bn_backward_reduce / maxpool_backward on bf16 rows of whole 16-byte words take the 128-bit forms (eight
    channels per thread); a buffer 2 bytes off a 16-byte boundary forces the scalar ones.  Same bits: per channel the
    arithmetic (fp32 strips of 64 rows into fp64 sums; the windows a voxel won, highest first) is the same."""
    from neraf_b200.gridnet import Window3d
    g = torch.Generator().manual_seed(13)
    V, c = 333, 24

    def off(rows, cols=c):              # bf16 view that starts 2 bytes past a 16-byte boundary, row stride cols + 8
        t = torch.empty(rows * (cols + 8) + 1, dtype=torch.bfloat16)[1:].view(rows, cols + 8)[:, :cols]
        assert t.data_ptr() % 16 != 0
        return t

    x = (torch.randn(V, c, generator=g) * 2).bfloat16()
    x_s = off(V); x_s.copy_(x)
    s_v, s_s = torch.empty(2, c, dtype=torch.float64), torch.empty(2, c, dtype=torch.float64)
    host_ops.bn_stats(x, s_v)
    host_ops.bn_stats(x_s, s_s)
    assert torch.equal(s_v, s_s)
    dy, dy2 = torch.randn(V, c, generator=g).bfloat16(), torch.randn(V, c, generator=g).bfloat16()
    y = torch.relu(torch.randn(V, c, generator=g)).bfloat16()
    mean, invstd = 0.1 * torch.randn(c, generator=g), 0.5 + torch.rand(c, generator=g)
    for second, gate in ((dy2, y), (None, y), (None, None)):
        g_v, g_s = torch.empty(V, c, dtype=torch.bfloat16), off(V)
        host_ops.bn_backward_reduce(dy, second, gate, x, mean, invstd, g_v, s_v)
        dy_s, x_o = off(V), off(V)
        dy_s.copy_(dy); x_o.copy_(x)
        sec_s = gate_s = None
        if second is not None:
            sec_s = off(V); sec_s.copy_(second)
        if gate is not None:
            gate_s = off(V); gate_s.copy_(gate)
        host_ops.bn_backward_reduce(dy_s, sec_s, gate_s, x_o, mean, invstd, g_s, s_s)
        assert torch.equal(g_v, g_s) and torch.equal(s_v, s_s)
    # pooling gradient on a (6, 5, 7) grid, 16 channels, with ties (post-ReLU zeros)
    w = Window3d(6, 5, 7, 16, 3, 2, 1)
    n_in, od, oh, ow = 6 * 5 * 7, 3, 3, 4
    xin = torch.relu(torch.randn(n_in, 16, generator=g)).bfloat16()
    yout, arg = torch.empty(od * oh * ow, 16, dtype=torch.bfloat16), torch.empty(od * oh * ow, 16, dtype=torch.int32)
    host_ops.maxpool(w, xin, yout, arg)
    gy, gy2 = torch.randn(od * oh * ow, 16, generator=g).bfloat16(), torch.randn(od * oh * ow, 16, generator=g).bfloat16()
    for second in (gy2, None):
        d_v, d_s = torch.empty(n_in, 16, dtype=torch.bfloat16), off(n_in, 16)
        host_ops.maxpool_backward(w, gy, second, arg, d_v)
        host_ops.maxpool_backward(w, gy, second, arg, d_s)
        assert torch.equal(d_v, d_s)


def test_reset_grid_is_the_reference_grid():
    """NeRAF_model.py:269-277 restated independently: zeros, voxel-centre coordinates in the last three channels."""
    from neraf_b200.model import NeRAFAudioModel, NeRAFAudioModelConfig
    for step in (1 / 64, 1 / 128):
        model = NeRAFAudioModel(NeRAFAudioModelConfig(dataset="RAF", grid_step=step), syn.default_aabb())
        assert model.grid is None
        model.reset_grid(device="cpu")
        n = int(1 / step)
        assert model.grid.shape == (7, n, n, n) and model.grid.dtype == torch.float32
        assert torch.count_nonzero(model.grid[:4]) == 0
        c = (torch.arange(n, dtype=torch.float64) + 0.5) * step
        assert torch.allclose(model.grid[4, :, 0, 0].double(), c, atol=1e-7)
        assert torch.allclose(model.grid[5, 3, :, 7].double(), c, atol=1e-7)
        assert torch.allclose(model.grid[6, 1, 2, :].double(), c, atol=1e-7)
        assert torch.all(model.grid[4, 5] == model.grid[4, 5, 0, 0]) and torch.all(model.grid[6, :, :, 9] == model.grid[6, 0, 0, 9])


def test_grid_cursor_and_cell_writer_follow_the_reference():
    """NeRAF_model.py:306-311,372-404 restated with a plain loop: the cursor visits every voxel centre once per sweep, a
    batch of field outputs lands in the cells its coordinates fall in, points outside the unit cube are dropped."""
    from neraf_b200.model import NeRAFAudioModel, NeRAFAudioModelConfig
    step = 1 / 16
    model = NeRAFAudioModel(NeRAFAudioModelConfig(dataset="RAF", grid_step=step), syn.default_aabb())
    model.reset_grid(device="cpu")
    seen = torch.cat([model.next_grid_batch(1000) for _ in range(5)])          # 16^3 = 4096 = 4 * 1000 + 96
    assert seen.shape == (4096, 3) and model.grid_batch_i == 0
    assert torch.equal(seen, model.grid[4:].permute(1, 2, 3, 0).reshape(-1, 3))
    assert torch.equal(model.next_grid_batch(10), seen[:10])                    # wrapped around
    g = torch.Generator().manual_seed(3)
    coords = torch.rand(500, 3, generator=g) * 1.2 - 0.1                        # some outside
    rgb, density = torch.randn(500, 3, generator=g), torch.rand(500, 1, generator=g) * 300
    before = model.grid.clone()
    model.write_grid_cells(coords, rgb, density)
    want = before.clone()
    for p in range(500):
        i, j, k = (int(coords[p, a] / step) for a in range(3))                  # .int() truncates toward zero
        if not all(0 <= v < 16 for v in (i, j, k)):
            continue
        want[:3, i, j, k] = torch.sigmoid(rgb[p])
        want[3, i, j, k] = torch.clip(1 - torch.exp(-1e-2 * density[p, 0]), 0, 1)
    # cells hit twice keep the last write in both versions only if the order is the same: compare cells hit once
    idx = torch.stack([(coords[:, a] / step).int() for a in range(3)], 1)
    ok = ((idx >= 0) & (idx < 16)).all(1)
    flat = (idx[ok, 0] * 256 + idx[ok, 1] * 16 + idx[ok, 2]).long()
    once = torch.bincount(flat, minlength=4096) == 1
    sel = once.view(16, 16, 16)
    assert torch.allclose(model.grid[:4][:, sel], want[:4][:, sel], atol=1e-7)
    untouched = torch.bincount(flat, minlength=4096).view(16, 16, 16) == 0
    assert torch.equal(model.grid[:4][:, untouched], before[:4][:, untouched])
    assert torch.equal(model.grid[4:], before[4:])
    model.write_grid_cells(coords[:5], rgb[:5].sigmoid(), density[:5], rendered_rgb=True)
    i, j, k = (int(coords[0, a] / step) for a in range(3))
    if all(0 <= v < 16 for v in (i, j, k)):
        assert torch.allclose(model.grid[:3, i, j, k], rgb[0].sigmoid(), atol=1e-7)
