"""Data-parallel host logic on CPU: gloo, world size 2 (the N > 1 path of bench.py / neraf_b200.distributed).

The CUDA kernels are not involved (no GPU here): the ranks run the ORACLE arithmetic on their shard and exchange
exactly what the product exchanges -- the four partial sums of the spectral loss and the parameter gradients -- so the
test pins the claim of DESIGN.md section 8: N-rank loss and gradients equal the single-process values on the
concatenated batch.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neraf_b200 import synthetic as syn
from neraf_b200.distributed import GradientAllReduce, shard_batch, shard_range
from oracle import loss as oloss


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        B, C, F = 64, 2, 33
        batch = {"data": torch.randn(B, C, F) - 1.0, "time_query": torch.arange(B), "scalar": 3}
        pred_full = (batch["data"] + 0.3 * torch.randn(B, C, F)).double()
        mine = shard_batch(batch, rank, world)
        lo, hi = shard_range(B, rank, world)
        assert mine["data"].shape[0] == B // world and mine["scalar"] == 3
        assert torch.equal(mine["time_query"], torch.arange(lo, hi))
        pred = pred_full[lo:hi].clone().requires_grad_(True)

        # the loss exchange: all-reduce the four partial sums, then finalize with the GLOBAL element count
        s_num, s_den, s_sq, s_abs = oloss.loss_sums(pred, mine["data"])
        sums = torch.stack([s_num, s_den, s_sq, s_abs])
        total = sums.detach().clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
        n_total = pred_full.numel()
        sc = 1e-4 * torch.sqrt(total[0]) / torch.sqrt(total[1])
        mag = 1e-3 * total[2] / n_total
        ref = oloss.loss_dict(pred_full, batch["data"])
        assert abs(float(sc) - float(ref["audio_sc_loss"])) < 1e-12 * float(ref["audio_sc_loss"])
        assert abs(float(mag) - float(ref["audio_mag_loss"])) < 1e-12 * float(ref["audio_mag_loss"])

        # local gradient formed from the GLOBAL sums == the matching rows of the single-process gradient
        a = 1e-4 / (torch.sqrt(total[0]) * torch.sqrt(total[1]))
        ex, ey = torch.exp(pred.detach()), torch.exp(mine["data"].double())
        dpred = a * (ex - ey) * ex + 2e-3 * (pred.detach() - mine["data"].double()) / n_total
        gref = oloss.loss_grad(pred_full, batch["data"])
        assert torch.allclose(dpred, gref[lo:hi], rtol=1e-10, atol=1e-18)

        # the gradient exchange: per-rank parameter gradients are partial sums over the shard
        w = torch.nn.Parameter(torch.ones(C * F, 5, dtype=torch.float64))
        b = torch.nn.Parameter(torch.zeros(5, dtype=torch.float64))
        x = pred_full[lo:hi].reshape(hi - lo, -1)
        (x @ w + b).square().sum().backward()
        GradientAllReduce([w, b], None)()
        w1 = torch.ones(C * F, 5, dtype=torch.float64, requires_grad=True)
        b1 = torch.zeros(5, dtype=torch.float64, requires_grad=True)
        (pred_full.reshape(B, -1) @ w1 + b1).square().sum().backward()
        assert torch.allclose(w.grad, w1.grad, rtol=1e-12) and torch.allclose(b.grad, b1.grad, rtol=1e-12)
        avg = GradientAllReduce([w], None, average=True)
        w.grad = torch.full_like(w, float(rank + 1))
        avg()
        assert torch.allclose(w.grad, torch.full_like(w, (1 + world) * world / 2 / world))
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_two_rank_loss_and_gradient_exchange_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(tmp_path, f"ok{r}")) for r in range(world))


def test_shard_helpers():
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    with pytest.raises(ValueError):
        shard_batch({"x": torch.zeros(5, 2)}, 0, 2)
    b = syn.make_batch(syn.RAF, 8, seed=0)
    parts = [shard_batch(b, r, 2) for r in range(2)]
    assert torch.equal(torch.cat([p["mic_pose"] for p in parts]), b["mic_pose"])


def _producer_worker(rank: int, world: int, port: int, out_dir: str):
    """A replicated ResNet3D producer under data parallelism: dg is summed over the ranks on its way into the producer
    (model.grid_feature), so every rank's producer gradients equal the single-process gradients for the summed dg and
    the producer's parameters stay out of the gradient exchange."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from neraf_b200.gridnet import ResNet3D_helper
        from neraf_b200.model import NeRAFAudioModel, NeRAFAudioModelConfig
        from tests.gridnet_host import HostOps
        n = 64
        sd = syn.make_gridnet_state_dict("resnet18", seed=2)

        def make(group):
            net = ResNet3D_helper(7, "resnet18", False, 1 / n, 1024, precision="fp32")
            net.load_state_dict(sd)
            net.backbone_net.ops = HostOps()
            net.eval()
            cfg = NeRAFAudioModelConfig(dataset="RAF", grid_step=1 / n)
            return NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=net, grid=syn.make_grid(n)[0], process_group=group)

        model = make(dist.group.WORLD)
        probes = [torch.randn(256, generator=torch.Generator().manual_seed(40 + r)) for r in range(world)]
        feat = model.grid_feature()
        assert feat.shape == (256,) and feat.requires_grad
        (feat * probes[rank]).sum().backward()                     # this rank's shard of the loss
        single = make(None)                                        # one process, the whole batch
        (single.grid_feature() * sum(probes)).sum().backward()
        for (k, p), q in zip(model.resnet3d.named_parameters(), single.resnet3d.parameters()):
            assert torch.allclose(p.grad, q.grad, rtol=1e-4, atol=1e-6 * float(q.grad.abs().max())), k
        exchanged = model.data_parallel_parameters()
        assert not any(p is q for p in model.resnet3d.parameters() for q in exchanged)
        assert all(any(p is q for q in exchanged) for p in model.field.parameters())
        with torch.no_grad():                                      # no autograd graph: nothing to wrap, no collective
            assert not model.grid_feature().requires_grad
        open(os.path.join(out_dir, f"producer_ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_replicated_producer_receives_the_summed_feature_gradient(tmp_path, built):
    world = 2
    mp.spawn(_producer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"producer_ok{r}")) for r in range(world))
