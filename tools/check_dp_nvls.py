"""2+ GPU check (torchrun): the fused NVLS all-reduce (multimem.red in the weight-gradient GEMMs) gives the same summed
gradients as backward + one NCCL all-reduce, and both equal N x the single-rank gradients when every rank holds the
same batch."""
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
shape, B = syn.RAF, 512


def build():
    cfg = NeRAFAudioModelConfig(dataset="RAF", precision="bf16")
    m = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)),
                        process_group=dist.group.WORLD)
    m.field.load_state_dict(syn.make_state_dict(shape, seed=0))
    return m.to(dev)


batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, B, seed=10 + rank).items()}
res = {}
for name, fused in (("nccl", False), ("nvls", True)):
    model = build()
    step = GraphedTrainStep(model, batch, fused_allreduce=fused)
    for _ in range(3):
        ld = step(batch)
        step.allreduce_grads()
    torch.cuda.synchronize()
    res[name] = (torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.numel() > 0]).clone(), {k: float(v) for k, v in ld.items()}, step.nvls)
    dist.barrier()
a, b = res["nccl"][0], res["nvls"][0]
rel = float((a.double() - b.double()).norm() / a.double().norm())
mx = float((a - b).abs().max() / a.abs().max())
print(f"rank {rank}: nvls path active={res['nvls'][2]}  rel_fro(nvls, nccl)={rel:.3e}  max={mx:.3e}  "
      f"losses nccl={res['nccl'][1]} nvls={res['nvls'][1]}", flush=True)
ok = rel < 1e-5 and res["nvls"][2]
t = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if float(t) == 1.0 else 1)
