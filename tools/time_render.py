"""GPU tuning aid: pieces of NeRAFAudioModel.render_rirs for 1024 poses."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neraf_b200 import synthetic as syn
from neraf_b200.model import ConstantGridFeature, NeRAFAudioModel, NeRAFAudioModelConfig
dev = torch.device("cuda:0")
shape = syn.RAF
cfg = NeRAFAudioModelConfig(dataset=shape.name, max_len=shape.T, fs=shape.fs, N_freq_stft=shape.F, hop_len=shape.hop, win_len=shape.win, precision="bf16")
model = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)))
model.field.load_state_dict(syn.make_state_dict(shape, seed=0)); model = model.to(dev)
n = 1024
aabb = syn.default_aabb(); lo, hi = aabb[0] + 1.0, aabb[1] - 1.0
mic = (lo + (hi - lo) * torch.rand(n, 3)).double().pin_memory()
src = ((lo + hi) / 2).double().reshape(1, 3); rot = torch.tensor([[1.0, 0.5, 0.5]], dtype=torch.float64)
init = torch.rand(n, shape.C, shape.F, shape.T, dtype=torch.complex64, device=dev)
def t(fn, k=3):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
y = model.query_rirs(mic, src, rot)
print("query_rirs ms", t(lambda: model.query_rirs(mic, src, rot)))
print("gl.render(init) ms", t(lambda: model.istft_transform.render(y, init)))
print("gl.render(None) ms", t(lambda: model.istft_transform.render(y, None)))
print("render_rirs(init) ms", t(lambda: model.render_rirs(mic, src, rot, init)))
out = torch.empty(n, shape.C, shape.hop * (shape.T - 1)).pin_memory()
w = model.render_rirs(mic, src, rot, init)
print("d2h pinned ms", t(lambda: out.copy_(w, non_blocking=True)))
