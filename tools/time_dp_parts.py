"""1+ GPUs (torchrun for N > 1): where the data-parallel step's time goes.  Per variant: the graphed step, then the same
library calls issued eagerly with CUDA events between forward | backward (+ exchange beside it) | widening + grid-block
gradients.  Eager launches add host gaps; the SHARES are what this is for.  ONLY=... selects variants."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neraf_b200 import synthetic as syn  # noqa: E402
from neraf_b200.model import ConstantGridFeature, GraphedTrainStep, NeRAFAudioModel, NeRAFAudioModelConfig  # noqa: E402

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev, init_method=None if "RANK" in os.environ else "tcp://127.0.0.1:29533",
                        rank=rank, world_size=world)
shape, B = syn.RAF, int(os.environ.get("BATCH", "2048"))
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, B, seed=10 + rank).items()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def make(group, **kw):
    cfg = NeRAFAudioModelConfig(dataset="RAF", precision="bf16")
    model = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)),
                            process_group=group)
    model.field.load_state_dict(syn.make_state_dict(shape, seed=0))
    model = model.to(dev)
    return GraphedTrainStep(model, batch, **kw)


N_IT = int(os.environ.get("N_ITERS", "40"))


def timed(fn, n=None):
    n = n or N_IT
    tot = 0.0
    for it in range(n + 5):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        if it >= 5:
            tot += s.elapsed_time(e) * 1e3 / n
    t = torch.tensor([tot], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def parts(step, n=None):
    n = n or N_IT
    acc = [0.0, 0.0, 0.0]
    for it in range(n + 5):
        flush.fill_(1)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        for i, part in enumerate(step._parts):
            part()
            ev[i + 1].record()
        torch.cuda.synchronize()
        if it >= 5:
            for i in range(3):
                acc[i] += ev[i].elapsed_time(ev[i + 1]) * 1e3 / n
    t = torch.tensor(acc, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


only = os.environ.get("ONLY", "single,kernel").split(",")
if "single" in only:
    st = make(None)
    g = timed(lambda: st(batch))
    p = parts(st)
    if rank == 0:
        print(f"single process      graph {g:7.1f} us | eager parts: forward+loss sums {p[0]:6.1f}  backward {p[1]:6.1f}  (grid part {p[2]:5.1f})", flush=True)
    del st
if world > 1 or os.environ.get("ONE_RANK_GROUP"):
    for name, kw in (("kernel", dict(exchange="kernel")), ("nccl bf16", dict(exchange="nccl", grad_dtype=torch.bfloat16))):
        if name not in only:
            continue
        st = make(dist.group.WORLD, **kw)

        def whole():
            st(batch)
            st.allreduce_grads()
        g = timed(whole)
        line = f"{name:18s}  graph {g:7.1f} us"
        if st.kernel_exchange:
            p = parts(st)
            line += f" | eager parts: forward {p[0]:6.1f}  loss + backward + exchange {p[1]:6.1f}  widen + grid grads {p[2]:5.1f}"
        if rank == 0:
            print(line, flush=True)
        del st
dist.barrier()
dist.destroy_process_group()
