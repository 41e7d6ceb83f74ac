"""GPU parity tests of the acoustic field (forward, backward, loss) against the oracle and the golden
vectors produced by the reference's own classes.

Tolerances (BASELINE.json north_star; scale-relative as argued in SURVEY.md section 7):
  fp32 path: outputs 1e-5 (max-norm and Frobenius), losses 1e-5, parameter gradients 1e-4
  bf16 path: outputs 1e-2, losses 1e-2, parameter gradients 3e-2
"""
import os

import numpy as np
import pytest
import torch

from neraf_b200 import _lib
from neraf_b200 import synthetic as syn
from neraf_b200.field import NeRAFAudioSoundField
from neraf_b200.loss import spectral_loss
from oracle import encodings as oenc
from oracle import field as ofield
from oracle import loss as oloss
from tests.util import cuda, rel_fro, rel_max

pytestmark = pytest.mark.gpu

TOL = {"fp32": dict(y=1e-5, loss=1e-5, grad=1e-4), "bf16": dict(y=1e-2, loss=1e-2, grad=3e-2)}


def _bottom_tol(prec, rows_averaged):
    """Gradients at the bottom of the 6-layer chain (d/d grid feature, d/d h).  In bf16 every activation and
    activation-gradient is rounded to 8 bits of mantissa and LeakyReLU gates near zero can flip; summing over a
    batch averages that noise out (B >= 128 meets the 3e-2 gate), a per-row gradient or a tiny batch does not."""
    if prec == "fp32":
        return TOL["fp32"]["grad"]
    return TOL["bf16"]["grad"] if rows_averaged >= 128 else 1.5e-1


def _make_field(shape, sd, prec, dev):
    f = NeRAFAudioSoundField(1187, 512, sound_rez=shape.C, N_frequencies=shape.F, precision=prec)
    f.load_state_dict(sd)
    return f.to(dev)


def _oracle_step(shape, sd, batch, g):
    enc = oenc.encode_queries(batch, syn.default_aabb(), shape.T)
    y, acts = ofield.field_forward_factored(sd, enc, g, torch.float64, keep=True)
    ld = oloss.loss_dict(y, batch["data"])
    dy = oloss.loss_grad(y, batch["data"])
    grads, dgrid = ofield.field_backward(sd, acts, g, y, dy)
    return y, ld, dy, grads, dgrid


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("shape,B", [(syn.RAF, 200), (syn.SOUNDSPACES, 136), (syn.RAF, 1), (syn.RAF, 257), (syn.SOUNDSPACES, 513)])
def test_train_step_matches_oracle(prec, shape, B):
    dev = cuda()
    sd = syn.make_state_dict(shape, seed=2)
    batch = syn.make_batch(shape, B, seed=2, outside_frac=0.02)
    g = syn.make_grid_feature(2)
    y_ref, ld_ref, dy_ref, grads_ref, dgrid_ref = _oracle_step(shape, sd, batch, g)

    field = _make_field(shape, sd, prec, dev)
    gd = g.to(dev).requires_grad_(True)
    y = field.forward_queries(batch["time_query"], batch["mic_pose"], batch["source_pose"], batch["rot"],
                              syn.default_aabb().to(dev), shape.T, gd)
    assert y.shape == (B, shape.C, shape.F) and y.dtype == torch.float32
    sc, mag = spectral_loss(y, batch["data"].to(dev), "SC+SLMSE", 0.1 * 1e-3, 1e-3)
    (sc + mag).backward()
    t = TOL[prec]
    assert rel_fro(y, y_ref) < t["y"] and rel_max(y, y_ref) < t["y"]
    assert abs(float(sc) - float(ld_ref["audio_sc_loss"])) < t["loss"] * float(ld_ref["audio_sc_loss"])
    assert abs(float(mag) - float(ld_ref["audio_mag_loss"])) < t["loss"] * float(ld_ref["audio_mag_loss"])
    assert rel_fro(gd.grad, dgrid_ref) < _bottom_tol(prec, B)
    for name, p in field.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, name
        assert rel_fro(p.grad, grads_ref[name]) < _bottom_tol(prec, B), name


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["RAF", "SoundSpaces"])
def test_against_reference_golden(golden_dir, prec, name):
    """Same weights / inputs as oracle/make_golden.py fed to the reference's own NeRAFAudioSoundField + STFTLoss."""
    dev = cuda()
    gold = np.load(os.path.join(golden_dir, f"field_{name}.npz"))
    B, seed, C, F, T = [int(v) for v in gold["meta"]]
    shape = syn.RAF if name == "RAF" else syn.SOUNDSPACES
    sd = syn.make_state_dict(shape, seed=seed)
    batch = syn.make_batch(shape, B, seed=seed)
    g = syn.make_grid_feature(seed)
    field = _make_field(shape, sd, prec, dev)
    gd = g.to(dev).requires_grad_(True)
    y = field.forward_queries(batch["time_query"], batch["mic_pose"], batch["source_pose"], batch["rot"],
                              syn.default_aabb().to(dev), T, gd)
    sc, mag = spectral_loss(y, batch["data"].to(dev), "SC+SLMSE", 0.1 * 1e-3, 1e-3)
    (sc + mag).backward()
    t = TOL[prec]
    assert rel_fro(y, gold["y_f64"]) < t["y"] and rel_max(y, gold["y_f64"]) < t["y"]
    assert abs(float(sc) / 1e-4 - float(gold["sc_f64"])) < t["loss"] * float(gold["sc_f64"])
    assert abs(float(mag) / 1e-3 - float(gold["mag_f64"])) < t["loss"] * float(gold["mag_f64"])
    assert rel_fro(gd.grad, gold["dgrid_f64"]) < _bottom_tol(prec, B)
    gt = _bottom_tol(prec, B)                     # golden batches are small (24 / 16 rows): see _bottom_tol
    for pname, p in field.named_parameters():
        gn = float(gold[f"gnorm_f64:{pname}"])
        assert abs(float(p.grad.double().norm()) - gn) < gt * gn, pname
        if p.dim() == 2:
            assert rel_fro(p.grad[:4, -8:], gold[f"gslice_f64:{pname}"]) < 3 * gt, pname
        else:
            assert rel_fro(p.grad[:16], gold[f"gslice_f64:{pname}"]) < 3 * gt, pname


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_dense_forward_is_reference_signature(prec):
    """field.forward(h) on an arbitrary (B, 1187) input, incl. d/dh (NeRAF_field.py:47-65)."""
    dev = cuda()
    shape = syn.SOUNDSPACES
    sd = syn.make_state_dict(shape, seed=4)
    field = _make_field(shape, sd, prec, dev)
    g = torch.Generator().manual_seed(4)
    h = torch.randn(50, 1187, generator=g) * 0.5
    hd = h.to(dev).requires_grad_(True)
    y = field(hd)
    w = torch.randn(50, shape.C, shape.F, generator=g)
    (y * w.to(dev)).sum().backward()
    h64 = h.double().requires_grad_(True)
    y_ref = ofield.field_forward(sd, h64, torch.float64)
    (y_ref * w.double()).sum().backward()
    t = TOL[prec]
    assert rel_fro(y, y_ref) < t["y"]
    assert rel_fro(hd.grad, h64.grad) < _bottom_tol(prec, 1)


def test_no_grid_variant_uses_other_column_order():
    dev = cuda()
    shape = syn.RAF
    f = NeRAFAudioSoundField(163, 512, sound_rez=1, N_frequencies=513, precision="fp32").to(dev)
    batch = syn.make_batch(shape, 40, seed=9)
    y = f.forward_queries(batch["time_query"], batch["mic_pose"], batch["source_pose"], batch["rot"],
                          syn.default_aabb().to(dev), shape.T, None, order=_lib.ORDER_MIC_SRC_TIME_ROT)
    h = oenc.assemble_input(batch, syn.default_aabb(), shape.T, None)
    sd = {k: v.detach().cpu() for k, v in f.state_dict().items()}
    assert rel_fro(y, ofield.field_forward(sd, h, torch.float64)) < 1e-5


def test_inference_path_and_pack_refresh():
    """no_grad forward (keep=0) equals the training forward; an in-place parameter update invalidates the bf16 pack."""
    dev = cuda()
    shape = syn.RAF
    sd = syn.make_state_dict(shape, seed=6)
    field = _make_field(shape, sd, "bf16", dev)
    batch = syn.make_batch(shape, 64, seed=6)
    g = syn.make_grid_feature(6).to(dev)
    args = (batch["time_query"], batch["mic_pose"], batch["source_pose"], batch["rot"], syn.default_aabb().to(dev), shape.T, g)
    y_train = field.forward_queries(*args)
    with torch.no_grad():
        y_inf = field.forward_queries(*args)
        assert torch.equal(y_train.detach(), y_inf)
        field.STFT_linear[0].bias.add_(1.0)
        y_new = field.forward_queries(*args)
    assert float((y_new - y_inf).abs().max()) > 1e-3


def test_full_size_batch_properties():
    """BASELINE size (B=2048): row independence -- any slice of the batch gives the same rows; finite grads."""
    dev = cuda()
    shape = syn.RAF
    sd = syn.make_state_dict(shape, seed=0)
    field = _make_field(shape, sd, "bf16", dev)
    batch = syn.make_batch(shape, 2048, seed=0)
    g = syn.make_grid_feature(0).to(dev).requires_grad_(True)
    aabb = syn.default_aabb().to(dev)
    y = field.forward_queries(batch["time_query"], batch["mic_pose"], batch["source_pose"], batch["rot"], aabb, shape.T, g)
    sl = slice(777, 777 + 130)
    with torch.no_grad():
        y_sl = field.forward_queries(batch["time_query"][sl], batch["mic_pose"][sl], batch["source_pose"][sl],
                                     batch["rot"][sl], aabb, shape.T, g)
    assert torch.equal(y[sl].detach(), y_sl)
    sc, mag = spectral_loss(y, batch["data"].to(dev), "SC+SLMSE", 1e-4, 1e-3)
    (sc + mag).backward()
    for p in field.parameters():
        assert torch.isfinite(p.grad).all()
    assert float(y.abs().max()) <= 10.0


@pytest.mark.parametrize("shape", [syn.RAF, syn.SOUNDSPACES], ids=["RAF", "SoundSpaces"])
def test_full_size_batch_matches_oracle(shape):
    """BASELINE size (B = 2048 columns, the batch of NeRAF_config.py:47) against the oracle on the same inputs: the whole
    train step of the bf16 tensor-core path, outputs 1e-2 / gradients 3e-2 (north_star's bf16 tolerance), the float64
    oracle on the host (about a second at this size)."""
    dev = cuda()
    B = 2048
    sd = syn.make_state_dict(shape, seed=4)
    batch = syn.make_batch(shape, B, seed=4, outside_frac=0.02)
    g = syn.make_grid_feature(4)
    y_ref, ld_ref, _, grads_ref, dgrid_ref = _oracle_step(shape, sd, batch, g)
    field = _make_field(shape, sd, "bf16", dev)
    gd = g.to(dev).requires_grad_(True)
    y = field.forward_queries(batch["time_query"], batch["mic_pose"], batch["source_pose"], batch["rot"],
                              syn.default_aabb().to(dev), shape.T, gd)
    sc, mag = spectral_loss(y, batch["data"].to(dev), "SC+SLMSE", 0.1 * 1e-3, 1e-3)
    (sc + mag).backward()
    t = TOL["bf16"]
    assert rel_fro(y, y_ref) < t["y"] and rel_max(y, y_ref) < t["y"]
    assert abs(float(sc) - float(ld_ref["audio_sc_loss"])) < t["loss"] * float(ld_ref["audio_sc_loss"])
    assert abs(float(mag) - float(ld_ref["audio_mag_loss"])) < t["loss"] * float(ld_ref["audio_mag_loss"])
    assert rel_fro(gd.grad, dgrid_ref) < t["grad"]
    for name, p in field.named_parameters():
        assert rel_fro(p.grad, grads_ref[name]) < t["grad"], name


@pytest.mark.parametrize("policy", ["cp", "rb"])
@pytest.mark.parametrize("B", [2048, 5000])
def test_planned_tile_order_equals_static_stride(policy, B, monkeypatch):
    """The job-list kernel under an explicit tile plan (csrc/mega_plan.h: which CTA pair runs which tile, in which
    order) computes what it computes under the static stride: outputs and weight gradients bit for bit (a tile's
    arithmetic does not depend on where or when it runs), bias gradients up to the order of their atomics."""
    dev = cuda()
    shape = syn.RAF
    sd = syn.make_state_dict(shape, seed=1)
    batch = syn.make_batch(shape, B, seed=1)
    target = batch["data"].to(dev)
    aabb = syn.default_aabb().to(dev)
    res = {}
    for mode in ("static", policy):
        monkeypatch.setenv("NERAF_MEGA_PLAN", mode)
        field = _make_field(shape, sd, "bf16", dev)
        g = syn.make_grid_feature(1).to(dev).requires_grad_(True)
        y = field.forward_queries(batch["time_query"], batch["mic_pose"], batch["source_pose"], batch["rot"], aabb, shape.T, g)
        sc, mag = spectral_loss(y, target, "SC+SLMSE", 1e-4, 1e-3)
        (sc + mag).backward()
        torch.cuda.synchronize()
        res[mode] = (y.detach().clone(), {n: p.grad.clone() for n, p in field.named_parameters()}, g.grad.clone())
    y0, g0, dg0 = res["static"]
    y1, g1, dg1 = res[policy]
    assert torch.equal(y0, y1)
    for n in g0:
        if n.endswith("weight") and not n.startswith("soundfield.0."):
            assert torch.equal(g0[n], g1[n]), n
        else:                                   # bias gradients (atomics) and what is derived from db1
            assert rel_fro(g1[n], g0[n].cpu()) < 1e-5, n
    assert rel_fro(dg1, dg0.cpu()) < 1e-5


def test_cpu_parameters_fail_loudly():
    f = NeRAFAudioSoundField(1187, 512, 1, 513)
    with pytest.raises(_lib.NerafError):
        f(torch.zeros(2, 1187))


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_graphed_step_and_prefetch_equal_autograd_step(prec):
    """One-graph GraphedTrainStep (N = 1) fed pinned host batches -- directly and through prefetch() -- reproduces the
    eager plugin calls, batch after batch."""
    from neraf_b200.model import ConstantGridFeature, GraphedTrainStep, NeRAFAudioModel, NeRAFAudioModelConfig
    dev = cuda()
    shape, B = syn.RAF, 320
    cfg = NeRAFAudioModelConfig(dataset="RAF", precision=prec)
    model = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)))
    model.field.load_state_dict(syn.make_state_dict(shape, seed=0))
    model = model.to(dev)
    model.field.always_repack = True
    params = [p for p in model.parameters() if p.requires_grad and p.numel() > 0]
    host = [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, B, seed=s).items()}
            for s in (1, 2, 3)]

    def eager(batch):
        for p in params:
            p.grad = None
        ld = model.get_loss_dict(model.get_outputs(batch), batch)
        sum(ld.values()).backward()
        return {k: float(v) for k, v in ld.items()}, [p.grad.clone() for p in params]

    refs = [eager(b) for b in host]
    step = GraphedTrainStep(model, host[0], functional=False)      # the autograd calls, captured
    assert step.launches_per_step > 0
    for i in (2, 0):
        got = step(host[i])
        torch.cuda.synchronize()
        for k, v in refs[i][0].items():
            assert abs(float(got[k]) - v) < 1e-5 * abs(v), (k, i)
        for p, r in zip(params, refs[i][1]):
            assert rel_fro(p.grad, r) < 1e-4, i        # same kernels; the bias sums' atomics re-order
    step = GraphedTrainStep(model, host[0])                        # direct library calls, fused loss gradient
    assert 0 < step.launches_per_step

    def check(got, i):
        torch.cuda.synchronize()
        for k, v in refs[i][0].items():
            assert abs(float(got[k]) - v) < 1e-5 * abs(v), (k, i)
        assert float(step.total_loss) == pytest.approx(sum(refs[i][0].values()), rel=1e-5)
        for p, r in zip(params, refs[i][1]):
            assert rel_fro(p.grad, r) < 1e-4, i        # same kernels; the bias sums' atomics re-order

    for i in (1, 0, 2):
        check(step(host[i]), i)
    # the partial sums formed by the heads' epilogue (optional variant) == the loss kernel's on the same prediction
    step = GraphedTrainStep(model, host[0], fuse_loss_sums=True)
    assert step.launches_per_step > 0
    for i in (0, 2):
        check(step(host[i]), i)
    pred = step._keep[2]
    ref_sums = torch.zeros(5, dtype=torch.float64, device=dev)
    _lib.check(_lib.lib().neraf_spectral_loss_sums(pred.data_ptr(), step.static["data"].data_ptr(), pred.numel(),
                                                   ref_sums.data_ptr(), 0, _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    assert torch.allclose(step.sums[:4], ref_sums[:4], rtol=1e-9, atol=0.0)
    step.prefetch(host[1])
    for i in (1, 2, 0):                        # consume the staged batch, stage the next one under the step
        got = step(host[i])
        step.prefetch(host[(i + 1) % 3])
        check(got, i)
    check(step(host[2]), 2)                    # a batch that was NOT the staged one is copied directly
    check(step(host[1]), 1)                    # ... and the batch still staged is consumed afterwards


def test_two_graph_data_parallel_step_equals_autograd_step():
    """GraphedTrainStep with a process group (graphs of direct C-ABI calls, loss sums all-reduced between them, every
    gradient-exchange variant of neraf_field_backward_dp) == the eager autograd step, on a one-rank NCCL group."""
    import socket
    import torch.distributed as dist
    from neraf_b200.model import ConstantGridFeature, GraphedTrainStep, NeRAFAudioModel, NeRAFAudioModelConfig
    dev = cuda()
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    try:
        shape, B = syn.RAF, 384
        cfg = NeRAFAudioModelConfig(dataset="RAF", precision="bf16")
        model = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)),
                                process_group=dist.group.WORLD)
        model.field.load_state_dict(syn.make_state_dict(shape, seed=0))
        model = model.to(dev)
        batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, B, seed=1).items()}
        params = [p for p in model.parameters() if p.requires_grad and p.numel() > 0]
        out = model.get_outputs(batch)
        ld = model.get_loss_dict(out, batch)
        sum(ld.values()).backward()
        ref = [p.grad.clone() for p in params]
        ref_loss = {k: float(v) for k, v in ld.items()}
        # serial exchange (compact dW1 block, deferred grid-block gradients), two-phase backward with the bulk of the
        # exchange on a communication stream, bf16 exchange: all must reproduce the autograd step
        # ... and so must the ONE-graph step whose exchanges are the library's own kernels (exchange="kernel": the four
        # loss sums traded inside the fused loss kernel, the gradients by neraf_dp_exchange_grads beside the backward;
        # on one rank the protocol runs against the local buffers -- flags, parities, completion counters and all)
        for kw, tol in ((dict(exchange="nccl"), 1e-5), (dict(exchange="nccl", overlap_allreduce=True), 1e-5),
                        (dict(exchange="nccl", grad_dtype=torch.bfloat16), 4e-3), (dict(fused_allreduce=True), 1e-5),
                        (dict(exchange="kernel"), 4e-3), (dict(), 4e-3), (dict(exchange="kernel", _pull=True), 4e-3)):
            # default: the push form + widening launch; NERAF_EXCHANGE_PULL=1: the pull form (sums delivered as fp32 into
            # the .grad buffers by the exchange itself)
            kw = dict(kw)
            os.environ["NERAF_EXCHANGE_PULL"] = "1" if kw.pop("_pull", False) else "0"
            step = GraphedTrainStep(model, batch, **kw)
            os.environ.pop("NERAF_EXCHANGE_PULL")
            assert step.kernel_exchange == (kw.get("exchange", "auto") != "nccl" and not kw.get("fused_allreduce"))
            for _ in range(3):
                got = step(batch)
                step.allreduce_grads()
            torch.cuda.synchronize()
            for k in ref_loss:
                assert abs(float(got[k]) - ref_loss[k]) < 1e-5 * abs(ref_loss[k]), (k, kw)
            for p, r in zip(params, ref):
                assert rel_fro(p.grad, r) < tol, kw          # same kernels: only atomics order / bf16 rounding differ
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_empty_batch(prec):
    """B = 0: an empty (0, C, F) result, no launch, no error (the reference returns an empty tensor as well)."""
    dev = cuda()
    shape = syn.RAF
    field = _make_field(shape, syn.make_state_dict(shape, seed=0), prec, dev)
    z3 = torch.zeros(0, 3, dtype=torch.float64)
    y = field.forward_queries(torch.zeros(0, dtype=torch.int64), z3, z3, z3, syn.default_aabb().to(dev), shape.T,
                              syn.make_grid_feature(0).to(dev))
    assert y.shape == (0, shape.C, shape.F)
    h = field(torch.zeros(0, 1187, device=dev))
    assert h.shape == (0, shape.C, shape.F)


def test_autocast_gradscaler_step_and_non_finite_skip():
    """The reference trains under fp16 autocast with a GradScaler (NeRAF_config.py:79, nerfstudio's Trainer:
    ``scaler.scale(loss).backward(); scaler.step(opt); scaler.update()``).  The plugin's autograd nodes must (1) carry the
    scale through -- unscaled gradients equal the gradients of the unscaled loss --, and (2) propagate a non-finite
    loss to non-finite gradients, so that the scaler SKIPS the optimizer step and backs its scale off."""
    from neraf_b200.model import ConstantGridFeature, NeRAFAudioModel, NeRAFAudioModelConfig
    dev = cuda()
    shape, B = syn.RAF, 256
    cfg = NeRAFAudioModelConfig(dataset="RAF", precision="bf16")
    model = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)))
    model.field.load_state_dict(syn.make_state_dict(shape, seed=0))
    model = model.to(dev)
    params = [p for p in model.parameters() if p.requires_grad and p.numel() > 0]
    batch = syn.make_batch(shape, B, seed=5)

    def step(b, scaler, opt):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16):
            out = model.get_outputs(b)
            loss = sum(model.get_loss_dict(out, b, {}).values())
        scaler.scale(loss).backward()
        return loss

    # (1) the scale is carried through both nodes
    opt = torch.optim.SGD(params, lr=0.0)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    loss = step(batch, scaler, opt)
    assert loss.dtype == torch.float32 and torch.isfinite(loss)
    scaled = [p.grad.clone() for p in params]
    opt.zero_grad(set_to_none=True)
    out = model.get_outputs(batch)
    sum(model.get_loss_dict(out, batch, {}).values()).backward()
    for s, p in zip(scaled, params):
        assert rel_fro(s / 1024.0, p.grad) < 1e-3          # bias gradients: atomics order; bf16 dz rounding at another scale
    # (2) a NaN target -> NaN loss -> non-finite gradients -> the step is skipped and the scale halves
    opt = torch.optim.SGD(params, lr=1.0)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    bad = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
    bad["data"][3, 0, 7] = float("nan")
    before = [p.detach().clone() for p in params]
    loss = step(bad, scaler, opt)
    assert not torch.isfinite(loss)
    assert any(not torch.isfinite(p.grad).all() for p in params)
    scaler.step(opt)
    scaler.update()
    assert all(torch.equal(b, p.detach()) for b, p in zip(before, params)), "the optimizer step must have been skipped"
    assert scaler.get_scale() == 512.0
    # ... and the next finite batch trains again
    loss = step(batch, scaler, opt)
    scaler.step(opt)
    scaler.update()
    assert torch.isfinite(loss) and not torch.equal(before[0], params[0].detach())
