"""GPU tuning aid: A/B of two GraphedTrainStep variants in one process (alternating replays, L2 flushed)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neraf_b200 import synthetic as syn
from neraf_b200.model import ConstantGridFeature, GraphedTrainStep, NeRAFAudioModel, NeRAFAudioModelConfig
dev = torch.device("cuda:0")
shape, B = syn.RAF, int(os.environ.get("BATCH", "2048"))
cfg = NeRAFAudioModelConfig(dataset=shape.name, max_len=shape.T, fs=shape.fs, N_freq_stft=shape.F, hop_len=shape.hop, win_len=shape.win, precision="bf16")
model = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)))
model.field.load_state_dict(syn.make_state_dict(shape, seed=0)); model = model.to(dev)
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, B, seed=0).items()}
def build(name, **env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return GraphedTrainStep(model, batch)
    finally:
        for k, v in old.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
# NERAF_PDL is read once per process by the library: run the script twice (NERAF_PDL=0 / unset) for that A/B
variants = {"one loss launch (sums + barrier + gradient)": build("fused", NERAF_FUSED_LOSS="1"),
            "loss_sums + head_backward launches": build("split", NERAF_FUSED_LOSS="0"),
            "sums in the heads' epilogue": GraphedTrainStep(model, batch, fuse_loss_sums=True)}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tot = {k: 0.0 for k in variants}
R = 150
for r in range(R + 10):
    for name, g in variants.items():
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g(g.static); e.record(); torch.cuda.synchronize()
        if r >= 10: tot[name] += s.elapsed_time(e)
for k, v in tot.items():
    print(f"{k:45s} {v / R * 1e3:8.1f} us/step  ({variants[k].launches_per_step} launches)")
