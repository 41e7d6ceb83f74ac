"""GPU tuning aid: the job-list kernel's tile plans (csrc/mega_plan.h) against the static stride, one graphed train step
per NERAF_MEGA_PLAN mode and batch size, alternating replays, L2 flushed.  BATCHES=2048,16384 MODES=static,auto,cp,rb"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neraf_b200 import synthetic as syn  # noqa: E402
from neraf_b200.model import ConstantGridFeature, GraphedTrainStep, NeRAFAudioModel, NeRAFAudioModelConfig  # noqa: E402

dev = torch.device("cuda:0")
shape = syn.RAF
batches = [int(b) for b in os.environ.get("BATCHES", "2048,4096,16384,65536").split(",")]
modes = os.environ.get("MODES", "static,auto,cp,rb").split(",")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for B in batches:
    cfg = NeRAFAudioModelConfig(dataset=shape.name, max_len=shape.T, fs=shape.fs, N_freq_stft=shape.F, hop_len=shape.hop,
                                win_len=shape.win, precision="bf16")
    model = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)))
    model.field.load_state_dict(syn.make_state_dict(shape, seed=0))
    model = model.to(dev)
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, B, seed=0).items()}
    steps = {}
    for m in modes:
        os.environ["NERAF_MEGA_PLAN"] = m
        steps[m] = GraphedTrainStep(model, batch)
    os.environ.pop("NERAF_MEGA_PLAN", None)
    R = max(10, min(100, (2048 * 100) // B))
    tot = {m: 0.0 for m in modes}
    for r in range(R + 3):
        for m, g in steps.items():
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); g(g.static); e.record(); torch.cuda.synchronize()
            if r >= 3:
                tot[m] += s.elapsed_time(e)
    base = tot[modes[0]] / R * 1e3
    print(f"B={B:6d}: " + "   ".join(f"{m} {tot[m] / R * 1e3:8.1f} us ({tot[m] / R * 1e3 / base:5.3f})" for m in modes)
          + f"   frac(best) {89.54e6 * B / (min(tot.values()) / R * 1e-3) / 1350.8e12:.3f}", flush=True)
    del steps, model
