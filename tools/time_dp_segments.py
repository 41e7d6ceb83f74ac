"""2+ GPU (torchrun): per-segment device time of the two-graph data-parallel step, NCCL vs fused NVLS all-reduce."""
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
dist.init_process_group("nccl", device_id=dev)
shape, B = syn.RAF, 2048
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, B, seed=10 + rank).items()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, fused in (("nccl", False), ("nvls", True)):
    cfg = NeRAFAudioModelConfig(dataset="RAF", precision="bf16")
    model = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)),
                            process_group=dist.group.WORLD)
    model.field.load_state_dict(syn.make_state_dict(shape, seed=0))
    model = model.to(dev)
    step = GraphedTrainStep(model, batch, fused_allreduce=fused)
    seg = {k: 0.0 for k in ("fwd", "sums_ar", "bwd", "grad_ar")}
    n = 40
    for it in range(n + 5):
        flush.fill_(1)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        step.graph_fwd.replay()
        ev[1].record()
        dist.all_reduce(step.sums[:4], group=step.group)
        ev[2].record()
        step.graph_bwd.replay()
        ev[3].record()
        step.allreduce_grads()
        ev[4].record()
        torch.cuda.synchronize()
        if it >= 5:
            for i, k in enumerate(seg):
                seg[k] += ev[i].elapsed_time(ev[i + 1]) * 1e3 / n
    if rank == 0:
        print(name, "nvls" if step.nvls else "plain", {k: round(v, 1) for k, v in seg.items()}, "total us", round(sum(seg.values()), 1), flush=True)
    dist.barrier()
dist.destroy_process_group()
