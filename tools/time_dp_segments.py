"""2+ GPU (torchrun): device time per data-parallel train step (two/three-graph step) for the gradient-exchange variants."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neraf_b200 import synthetic as syn  # noqa: E402
from neraf_b200.model import ConstantGridFeature, GraphedTrainStep, NeRAFAudioModel, NeRAFAudioModelConfig  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
os.environ.setdefault("NCCL_MAX_CTAS", "16")
dist.init_process_group("nccl", device_id=dev)
shape, B = syn.RAF, 2048
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, B, seed=10 + rank).items()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = None
variants = [("nccl fp32", dict(exchange="nccl", grad_dtype=torch.float32)),
            ("nccl bf16", dict(exchange="nccl", grad_dtype=torch.bfloat16)),
            ("kernel", dict(exchange="kernel"))]
if os.environ.get("ONLY"):
    variants = [v for v in variants if v[0] in os.environ["ONLY"].split(",")]
if os.environ.get("WITH_OVERLAP"):
    variants += [("overlap fp32", dict(exchange="nccl", overlap_allreduce=True, grad_dtype=torch.float32)),
                 ("overlap bf16", dict(exchange="nccl", overlap_allreduce=True, grad_dtype=torch.bfloat16))]
if os.environ.get("WITH_NVLS"):
    variants.append(("nvls fused", dict(fused_allreduce=True)))
for name, kw in variants:
    cfg = NeRAFAudioModelConfig(dataset="RAF", precision="bf16")
    model = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)),
                            process_group=dist.group.WORLD)
    model.field.load_state_dict(syn.make_state_dict(shape, seed=0))
    model = model.to(dev)
    step = GraphedTrainStep(model, batch, **kw)
    n, tot = 40, 0.0
    for it in range(n + 5):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step(batch)
        step.allreduce_grads()
        e.record()
        torch.cuda.synchronize()
        if it >= 5:
            tot += s.elapsed_time(e) * 1e3 / n
    g = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.numel() > 0]).double()
    if ref is None:
        ref = g
    err = float((g - ref).norm() / ref.norm())
    t = torch.tensor([tot], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{name:14s} {float(t):8.1f} us/step   kernel_exchange={step.kernel_exchange} "
              f"multicast={bool(step._xchg and step._xchg['multicast'])} overlap={getattr(step, 'overlap', False)} "
              f"nvls={step.nvls}   grads vs nccl fp32: {err:.2e}", flush=True)
    tr = getattr(step, "_comm_trace", None)
    if tr is not None:                     # timeline of the exchange kernel of the LAST step, this rank, us from its start
        t = tr.cpu().tolist()
        t0 = t[0]
        line = f"rank {rank} exchange kernel: workers done {(t[1] - t0) / 1e3:.1f}, every rank done {(t[2] - t0) / 1e3:.1f} us; chunks"
        for c, ch in enumerate(step._exchange_chunks):
            a, b, d = ((t[4 + 4 * c + i] - t0) / 1e3 for i in range(3))
            line += f" | {ch[1] / 1e6:.1f} MB: announced {a:.1f}, all ranks {b:.1f}, issued {d:.1f}"
        for r in range(world):
            if r == rank and r in (0, world - 1):
                print(line, flush=True)
            dist.barrier()
    dist.barrier()
dist.destroy_process_group()
