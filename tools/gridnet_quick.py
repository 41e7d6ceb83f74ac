"""One training-mode forward + backward of the grid-feature producer (ResNet3D-50, bf16) on a (1, 7, N, N, N) grid:
CUDA-event time per step, launches per step, finiteness, and the evaluation-mode linearity property.  Writes
gpurun_out/gridnet_<N>.json step by step (the file is valid after every stage)."""
import json
import os
import sys
import time

t_start = time.time()
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neraf_b200 import _lib, synthetic as syn          # noqa: E402
from neraf_b200.gridnet import ResNet3D_helper, conv_flops   # noqa: E402

_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(_pos[0]) if _pos else 128
prec = _pos[1] if len(_pos) > 1 else "bf16"
with_graph = "--graph" in sys.argv          # also capture the step in a CUDA graph and time its replay
out_path = os.path.join(ROOT, "gpurun_out", f"gridnet_{n}_{prec}.json")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
res = {"grid": [1, 7, n, n, n], "precision": prec, "import_s": time.time() - t_start}


def dump():
    with open(out_path, "w") as f:
        json.dump(res, f)


dev = torch.device("cuda:0")
lib = _lib.lib()
net = ResNet3D_helper(in_channels=7, backbone="resnet50", grid_step=1.0 / n, N_features=1024, precision=prec)
net.load_state_dict(syn.make_gridnet_state_dict("resnet50"))
net = net.to(dev).train()
grid = syn.make_grid(n).to(dev)
dfeat = torch.randn(1, 1024, 1, 1, 1, generator=torch.Generator().manual_seed(7)).to(dev)
res["setup_s"] = time.time() - t_start
dump()
for _ in range(2):
    net(grid).backward(dfeat)
torch.cuda.synchronize()
res["warm_s"] = time.time() - t_start
dump()
k = 5
l0 = lib.neraf_launch_count()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(k):
    net(grid).backward(dfeat)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / k
fl = conv_flops(net.backbone_net, n)
res.update({"ms_per_step": ms, "gflop_per_step": fl / 1e9, "achieved_tflops": fl / (ms * 1e-3) / 1e12,
            "launches_per_step": (lib.neraf_launch_count() - l0) // k,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9})
s.record()
with torch.no_grad():
    for _ in range(k):
        f = net(grid)
e.record()
torch.cuda.synchronize()
res["ms_forward_only"] = s.elapsed_time(e) / k
res["finite"] = bool(torch.isfinite(f).all() and all(torch.isfinite(p.grad).all() for p in net.parameters()))
dump()
print(json.dumps(res))
if with_graph:
    # the whole training step (forward + backward of every layer, ~440 launches) as ONE graph launch: what removes the
    # Python / ctypes launch sequence from the step.  Standard whole-network capture: warm up on a side stream, drop the
    # gradients so that the capture allocates them from the graph's pool, capture, replay.
    try:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                net.zero_grad(set_to_none=True)
                net(grid).backward(dfeat)
        torch.cuda.current_stream(dev).wait_stream(side)
        eager_grads = [p.grad.clone() for p in net.parameters()]
        net.zero_grad(set_to_none=True)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            feat_g = net(grid)
            feat_g.backward(dfeat)
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
        s.record()
        for _ in range(k):
            graph.replay()
        e.record()
        torch.cuda.synchronize()
        res["graph_ms_per_step"] = s.elapsed_time(e) / k
        res["graph_achieved_tflops"] = fl / (res["graph_ms_per_step"] * 1e-3) / 1e12
        # training-mode statistics are re-formed from the same grid and weights: the replay must reproduce the eager step
        res["graph_vs_eager_grad_max_rel"] = max(
            float((p.grad - a).norm() / a.norm().clamp_min(1e-30)) for p, a in zip(net.parameters(), eager_grads))
        del graph
    except Exception as exc:                                     # noqa: BLE001
        res["graph_error"] = f"{type(exc).__name__}: {exc}"
    dump()
    print(json.dumps(res))
net.eval()
net.zero_grad(set_to_none=True)
f1 = net(grid)
f1.backward(dfeat)
g1 = [p.grad.clone() for p in net.parameters()]
net.zero_grad(set_to_none=True)
f2 = net(grid)
f2.backward(2 * dfeat)
torch.cuda.synchronize()
res["eval_feature_repeatable"] = bool(torch.equal(f1, f2))
res["eval_backward_linear_max_rel"] = max(float((p.grad - 2 * a).norm() / (2 * a).norm().clamp_min(1e-30))
                                          for p, a in zip(net.parameters(), g1))
res["total_s"] = time.time() - t_start
dump()
print(json.dumps(res))
