"""Debug aid: eager step vs GraphedTrainStep(functional=True/False) with a real producer, per-parameter deviations."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neraf_b200 import synthetic as syn
from neraf_b200.model import GraphedTrainStep, NeRAFAudioModel, NeRAFAudioModelConfig
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
N, GRID_STEP = 64, 1 / 64
shape, B = syn.RAF, 256
cfg = NeRAFAudioModelConfig(dataset="RAF", precision=prec, grid_step=GRID_STEP, grid_net="resnet50")
model = NeRAFAudioModel(cfg, syn.default_aabb(), grid=syn.make_grid(N)[0])
model.field.load_state_dict(syn.make_state_dict(shape, seed=0))
model.resnet3d.load_state_dict(syn.make_gridnet_state_dict("resnet50"))
model = model.to(dev); model.grid = model.grid.to(dev)
model.resnet3d.eval(); model.field.always_repack = True
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, B, seed=1).items()}
params = [(n, p) for n, p in model.named_parameters() if p.requires_grad and p.numel() > 0]
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))
for _, p in params: p.grad = None
ld = model.get_loss_dict(model.get_outputs(batch), batch); sum(ld.values()).backward(); torch.cuda.synchronize()
ref = {n: p.grad.clone() for n, p in params}
ld = {k: float(v) for k, v in ld.items()}          # drop the autograd graph (its AccumulateGrad nodes sit on the legacy stream)
for functional in (True, False, False, True, False):
    step = GraphedTrainStep(model, batch, functional=functional)
    for _ in range(2): got = step(batch)
    torch.cuda.synchronize()
    errs = sorted(((rel(p.grad, ref[n]), n) for n, p in params), reverse=True)
    fld = max(e for e, n in errs if n.startswith("field"))
    prod = max(e for e, n in errs if not n.startswith("field"))
    print(f"functional={functional} fused_loss={getattr(step, '_fused_loss', None)} PDL={os.environ.get('NERAF_PDL')}: "
          f"field max {fld:.2e}  producer max {prod:.2e}  worst: {errs[0][1]} {errs[0][0]:.2e}; loss {[float(v) for v in got.values()]} vs {list(ld.values())}")
    print("   deviating:", [(n.replace("resnet3d.backbone_net.", ""), f"{e:.1e}") for e, n in errs if e > 1e-6][:12])
    # eager again: is the eager step itself repeatable?
    for _, p in params: p.grad = None
    ld2 = model.get_loss_dict(model.get_outputs(batch), batch); sum(ld2.values()).backward(); torch.cuda.synchronize(); del ld2
    e2 = sorted(((rel(p.grad, ref[n]), n) for n, p in params), reverse=True)
    print("   eager repeat: worst", e2[0][1], f"{e2[0][0]:.1e}")
