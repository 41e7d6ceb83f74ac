"""N-GPU check (torchrun): data-parallel training equals single-GPU training on the concatenated batch
(the reference refuses world_size > 1, so this equality IS the specification).

Every rank holds its own B columns; the DP step (two CUDA graphs, loss sums all-reduced in between, fused loss gradient,
compact dW1 block, deferred grid-block gradients, fp32 or bf16 gradient exchange) must give the loss and the SUMMED
gradients that one process computes with the eager plugin calls on all N*B columns."""
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
shape, B = syn.RAF, 768
KEYS = ("time_query", "mic_pose", "source_pose", "rot", "data")


def build(group):
    cfg = NeRAFAudioModelConfig(dataset="RAF", precision="bf16")
    m = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)),
                        process_group=group)
    m.field.load_state_dict(syn.make_state_dict(shape, seed=0))
    return m.to(dev)


def flat(model):
    return torch.cat([p.grad.reshape(-1).float() for p in model.parameters() if p.requires_grad and p.numel() > 0])


batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, B, seed=10 + rank).items()}
whole = {}
for k in KEYS:
    parts = [torch.empty_like(batch[k]) for _ in range(world)]
    dist.all_gather(parts, batch[k].contiguous())
    whole[k] = torch.cat(parts)

single = build(None)
ld = single.get_loss_dict(single.get_outputs(whole), whole)
sum(ld.values()).backward()
ref_g, ref_l = flat(single), {k: float(v) for k, v in ld.items()}

ok = True
for name, kw, tol in (("NCCL, fp32 exchange (two graphs)", dict(exchange="nccl", grad_dtype=torch.float32), 1e-4),
                      ("NCCL, bf16 exchange (two graphs)", dict(exchange="nccl", grad_dtype=torch.bfloat16), 4e-3),
                      ("kernel exchange (ONE graph: sums traded in the loss kernel, NVLS all-reduce beside the backward)",
                       dict(exchange="kernel"), 4e-3)):
    model = build(dist.group.WORLD)
    step = GraphedTrainStep(model, batch, **kw)
    for _ in range(4):
        got = step(batch)
        step.allreduce_grads()
    torch.cuda.synchronize()
    g = flat(model)
    # every rank must hold the SAME gradients bit for bit with the kernel exchange (one reduced buffer, deterministic dg)
    gmax, gmin = g.clone(), g.clone()
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX); dist.all_reduce(gmin, op=dist.ReduceOp.MIN)
    identical = bool((gmax == gmin).all())
    err = float((g - ref_g).norm() / ref_g.norm())
    lerr = max(abs(float(got[k]) - ref_l[k]) / abs(ref_l[k]) for k in ref_l)
    ok = ok and err < tol and lerr < 1e-5
    if rank == 0:
        print(f"{world} ranks x {B} columns, {name}: summed gradients vs single process on {world * B} columns: "
              f"rel {err:.2e} (tol {tol:g}); losses rel {lerr:.1e}; ranks bit-identical: {identical}"
              f"{' [multicast]' if getattr(step, '_xchg', None) and step._xchg['multicast'] else ''}", flush=True)
    dist.barrier()
if rank == 0:
    print("DP == single:", "OK" if ok else "MISMATCH", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
