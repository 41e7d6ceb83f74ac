"""The audio train step as the reference runs it (NeRAF_model.py:554-566): ResNet3D-50 producer on the (1, 7, N, N, N)
grid -> 1024-feature -> field -> spectral loss -> backward through the field AND the producer, captured as ONE CUDA
graph by GraphedTrainStep (training-mode batch norm).  Prints one JSON line: ms per step, columns/s.
usage: step_with_producer.py [N=128] [batch=2048]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neraf_b200 import synthetic as syn  # noqa: E402
from neraf_b200.model import GraphedTrainStep, NeRAFAudioModel, NeRAFAudioModelConfig  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
dev = torch.device("cuda:0")
shape = syn.RAF
cfg = NeRAFAudioModelConfig(dataset="RAF", precision="bf16", grid_step=1.0 / n, grid_net="resnet50")
model = NeRAFAudioModel(cfg, syn.default_aabb(), grid=syn.make_grid(n)[0])
model.field.load_state_dict(syn.make_state_dict(shape, seed=0))
model.resnet3d.load_state_dict(syn.make_gridnet_state_dict("resnet50"))
model = model.to(dev)
model.grid = model.grid.to(dev)
model.train()
model.field.always_repack = True
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, B, seed=0).items()}
step = GraphedTrainStep(model, batch)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
k, tot = 20, 0.0
for it in range(k + 3):
    flush.fill_(1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    step(step.static)
    e.record()
    torch.cuda.synchronize()
    if it >= 3:
        tot += s.elapsed_time(e) / k
finite = all(torch.isfinite(p.grad).all().item() for p in model.parameters() if p.grad is not None)
print(json.dumps({"ms_per_step": tot, "value": B / (tot * 1e-3), "unit": "columns/s", "batch": B, "grid": [1, 7, n, n, n],
                  "launches_per_step": step.launches_per_step, "finite": finite,
                  "api": "GraphedTrainStep over NeRAFAudioModel(grid_net='resnet50'): producer forward, field step, producer "
                         "backward in one CUDA graph, training-mode batch norm"}))
