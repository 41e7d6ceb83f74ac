"""Host-side profile of the eager plugin calls (get_outputs -> get_loss_dict -> backward -> loss.item()) on a pinned host
batch: where the ~0.3 ms per step of Python / autograd / ctypes goes (cProfile, top functions by cumulative time)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neraf_b200 import synthetic as syn  # noqa: E402
from neraf_b200.model import ConstantGridFeature, NeRAFAudioModel, NeRAFAudioModelConfig  # noqa: E402

dev = torch.device("cuda:0")
shape, B = syn.RAF, 2048
cfg = NeRAFAudioModelConfig(dataset="RAF", precision="bf16")
model = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)))
model.field.load_state_dict(syn.make_state_dict(shape, seed=0))
model = model.to(dev)
model.field.always_repack = True
params = [p for p in model.parameters() if p.requires_grad and p.numel() > 0]
host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, B, seed=0).items()}


def step():
    for p in params:
        p.grad = None
    out = model.get_outputs(host)
    ld = model.get_loss_dict(out, host)
    loss = sum(ld.values())
    loss.backward()
    return loss.item()


for _ in range(20):
    step()
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n):
    step()
torch.cuda.synchronize()
print(f"eager step: {(time.perf_counter() - t0) / n * 1e3:.3f} ms wall")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
