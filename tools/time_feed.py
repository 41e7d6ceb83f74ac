"""GPU tuning aid: times the pieces of a resident-feed training step."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neraf_b200 import synthetic as syn
from neraf_b200.datafeed import ResidentAudioFeed

dev = torch.device("cuda:0")
shape, B, n = syn.RAF, 2048, 2048
cache = (torch.randn(n, shape.T, shape.C, shape.F) - 3.0).to(dev)
pb = syn.make_batch(shape, n, seed=5)
feed = ResidentAudioFeed(cache, pb["mic_pose"], pb["source_pose"], pb["rot"], shape.T, B, seed=1, drop_last=True)
_, static = feed.next_train(0)

def t(fn, k=50):
    torch.cuda.synchronize(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter(); s.record()
    for _ in range(k): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / k * 1e3, (time.perf_counter() - w0) / k * 1e6

idx = feed._next_indices()
print("gather only (gpu us, wall us):", t(lambda: feed.batch_from_indices(idx, out=static)))
print("next_indices:", t(lambda: feed._next_indices()))
print("next_train:", t(lambda: feed.next_train(0, out=static)))

from neraf_b200.model import ConstantGridFeature, GraphedTrainStep, NeRAFAudioModel, NeRAFAudioModelConfig
cfg = NeRAFAudioModelConfig(dataset=shape.name, max_len=shape.T, fs=shape.fs, N_freq_stft=shape.F, hop_len=shape.hop,
                            win_len=shape.win, precision="bf16")
model = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)))
model.field.load_state_dict(syn.make_state_dict(shape, seed=0))
model = model.to(dev)
model.field.always_repack = True
dev_batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, B, seed=0).items()}
graphed = GraphedTrainStep(model, dev_batch)
print("graph on its own batch:", t(lambda: graphed(graphed.static)))
def both():
    feed.next_train(0, out=graphed.static)
    return graphed(graphed.static)
print("feed + graph:", t(both))
def both_item():
    feed.next_train(0, out=graphed.static)
    return sum(graphed(graphed.static).values()).item()
print("feed + graph + item:", t(both_item))
print("graph + item (own batch restored):", t(lambda: sum(graphed(dev_batch).values()).item()))

def v1():
    feed.batch_from_indices(idx, out=graphed.static)
    return sum(graphed(graphed.static).values()).item()
print("fixed-idx gather + graph + item:", t(v1))
def v2():
    feed._next_indices()
    return sum(graphed(graphed.static).values()).item()
print("next_indices + graph + item:", t(v2))
def v3():
    return sum(graphed(graphed.static).values()).item()
print("graph(static) + item:", t(v3))
def v4():
    feed.batch_from_indices(idx, out=graphed.static)
    graphed(graphed.static)
    torch.cuda.synchronize()
print("fixed-idx gather + graph + sync:", t(v4))
import time as _t
def v5():
    w0 = _t.perf_counter(); feed.batch_from_indices(idx, out=graphed.static); w1 = _t.perf_counter()
    ld = graphed(graphed.static); w2 = _t.perf_counter()
    x = sum(ld.values()); w3 = _t.perf_counter(); x.item(); w4 = _t.perf_counter()
    return (w1 - w0, w2 - w1, w3 - w2, w4 - w3)
torch.cuda.synchronize()
rs = [v5() for _ in range(20)]
print("wall us: gather %.1f replay %.1f sum %.1f item %.1f" % tuple(1e6 * sum(r[i] for r in rs[5:]) / 15 for i in range(4)))
