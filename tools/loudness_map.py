"""BASELINE config 4: the loudness-map workload of viz/loudness_maps.ipynb -- a 64 x 64 xz grid x 4
heights = 16 384 microphone poses, one fixed source / rotation, every pose rendered to a waveform (T field queries +
log->magnitude + Griffin-Lim) and reduced to its RMS loudness in dB.  Prints one JSON line."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neraf_b200 import synthetic as syn  # noqa: E402
from neraf_b200.model import ConstantGridFeature, NeRAFAudioModel, NeRAFAudioModelConfig  # noqa: E402

dev = torch.device("cuda:0")
shape = syn.RAF
cfg = NeRAFAudioModelConfig(dataset=shape.name, max_len=shape.T, fs=shape.fs, N_freq_stft=shape.F, hop_len=shape.hop,
                            win_len=shape.win, precision="bf16")
model = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)))
model.field.load_state_dict(syn.make_state_dict(shape, seed=0))
model = model.to(dev)
aabb = syn.default_aabb()
xs = torch.linspace(float(aabb[0, 0]) + 0.5, float(aabb[1, 0]) - 0.5, 64, dtype=torch.float64)
zs = torch.linspace(float(aabb[0, 2]) + 0.5, float(aabb[1, 2]) - 0.5, 64, dtype=torch.float64)
ys = torch.linspace(float(aabb[0, 1]) + 0.5, float(aabb[1, 1]) - 0.5, 4, dtype=torch.float64)
mic = torch.stack(torch.meshgrid(xs, ys, zs, indexing="ij"), -1).reshape(-1, 3).pin_memory()          # (16384, 3)
src = ((aabb[0] + aabb[1]) / 2).double().reshape(1, 3)
rot = torch.tensor([[1.0, 0.5, 0.5]], dtype=torch.float64)
chunk = 2048


def run():
    out = []
    for lo in range(0, mic.shape[0], chunk):
        w = model.render_rirs(mic[lo:lo + chunk], src, rot)                                           # (n, C, L) on the device
        out.append(10.0 * torch.log10(w.pow(2).mean(dim=(1, 2)) + 1e-12))                              # loudness (dB) per pose
    return torch.cat(out).cpu()


run()
torch.cuda.synchronize()
t0 = time.perf_counter()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
loud = run()
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e)
print(json.dumps({"workload": "loudness map: 64x64x4 = 16384 poses x T=60 queries -> Griffin-Lim -> RMS dB (RAF shapes, bf16)",
                  "poses": mic.shape[0], "queries": mic.shape[0] * shape.T, "ms": ms, "wall_ms": (time.perf_counter() - t0) * 1e3,
                  "rirs_per_sec": mic.shape[0] / (ms * 1e-3), "queries_per_sec": mic.shape[0] * shape.T / (ms * 1e-3),
                  "chunk": chunk, "loudness_db_range": [float(loud.min()), float(loud.max())]}))
