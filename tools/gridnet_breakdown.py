"""Where a training step of the grid-feature producer goes: every GridOps call of one forward + backward is timed
(CUDA events around each call, one synchronize at the end) and summed per operator, next to the bytes it moves, and the
step's wall time is compared with the sum of the device times (a large gap = the step is bound by the Python / ctypes
launch sequence, i.e. wants a CUDA graph).

    python tools/gridnet_breakdown.py [N=128] [bf16|fp32]          # on the B200
    python tools/gridnet_breakdown.py 64 fp32 --host               # dry run on the CPU (host build of the kernel code)

Writes gpurun_out/gridnet_breakdown_<N>_<prec>.json and prints a table."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neraf_b200 import synthetic as syn                     # noqa: E402
from neraf_b200.gridnet import GridOps, ResNet3D_helper, conv_flops      # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
host = "--host" in sys.argv
n = int(args[0]) if args else 128
prec = args[1] if len(args) > 1 else "bf16"

if host:
    from tests.gridnet_host import HostOps as Base
else:
    Base = GridOps

TIMED = ("im2col", "col2im", "pack_weight", "unpack_wgrad", "bn_stats", "bn_finalize", "bn_apply", "bn_backward_reduce",
         "bn_backward_apply", "maxpool", "maxpool_backward", "broadcast_rows", "gemm_nt", "gemm_backward")


def _bytes(ts):
    return sum(t.numel() * t.element_size() for t in ts if torch.is_tensor(t))


class TimedOps(Base):
    """Same calls, each bracketed by a pair of events (device) or perf_counter (host dry run)."""

    def __init__(self):
        super().__init__()
        self.records = []          # (name, start, stop, bytes, flops)
        self.enabled = False

    def _wrap(self, name):
        inner = getattr(Base, name)

        def call(*a, **k):
            if not self.enabled:
                return inner(self, *a, **k)
            nbytes = _bytes(a)
            flops = 0.0
            if name == "gemm_nt":                      # (A, B, M, N, K, out)
                flops = 2.0 * a[2] * a[3] * a[4]
            elif name == "gemm_backward":              # (dy, wmat, col, v_out, kc, c_out, dw_mat, dcol): wgrad (+ dgrad)
                flops = 2.0 * a[3] * a[4] * a[5] * (2 if a[7] is not None else 1)
            if host:
                t0 = time.perf_counter()
                inner(self, *a, **k)
                self.records.append((name, t0, time.perf_counter(), nbytes, flops))
            else:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                inner(self, *a, **k)
                e.record()
                self.records.append((name, s, e, nbytes, flops))
        return call

    def __getattribute__(self, item):
        if item in TIMED:
            return object.__getattribute__(self, "_wrap")(item)
        return object.__getattribute__(self, item)


dev = torch.device("cpu" if host else "cuda:0")
net = ResNet3D_helper(in_channels=7, backbone="resnet50", grid_step=1.0 / n, N_features=1024, precision=prec)
net.load_state_dict(syn.make_gridnet_state_dict("resnet50"))
net = net.to(dev).train()
ops = TimedOps()
net.backbone_net.ops = ops
grid = syn.make_grid(n).to(dev)
dfeat = torch.randn(1, 1024, 1, 1, 1, generator=torch.Generator().manual_seed(7)).to(dev)
sync = (lambda: None) if host else torch.cuda.synchronize
for _ in range(1 if host else 3):
    net(grid).backward(dfeat)
sync()
# (a) the step as it runs: wall clock, no per-call events
t0 = time.perf_counter()
reps = 1 if host else 5
for _ in range(reps):
    net(grid).backward(dfeat)
sync()
wall_ms = (time.perf_counter() - t0) * 1e3 / reps
# (b) the same step with every call bracketed
ops.enabled = True
net(grid).backward(dfeat)
sync()
table = {}
for name, a, b, nbytes, flops in ops.records:
    ms = (b - a) * 1e3 if host else a.elapsed_time(b)
    r = table.setdefault(name, {"calls": 0, "ms": 0.0, "bytes": 0, "flops": 0.0})
    r["calls"] += 1; r["ms"] += ms; r["bytes"] += nbytes; r["flops"] += flops
dev_ms = sum(r["ms"] for r in table.values())
res = {"grid": n, "precision": prec, "host_dry_run": host, "wall_ms_per_step": wall_ms, "sum_of_call_ms": dev_ms,
       "gflop_per_step": conv_flops(net.backbone_net, n) / 1e9, "ops": table}
out = os.path.join(ROOT, "gpurun_out", f"gridnet_breakdown_{n}_{prec}{'_host' if host else ''}.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(res, open(out, "w"), indent=1)
print(f"{'op':22s} {'calls':>5s} {'ms':>9s} {'GB (args)':>10s} {'GB/s':>8s} {'TFLOP/s':>8s}")
for name, r in sorted(table.items(), key=lambda kv: -kv[1]["ms"]):
    gbs = r["bytes"] / 1e9 / (r["ms"] * 1e-3) if r["ms"] > 0 else 0.0
    tf = r["flops"] / 1e12 / (r["ms"] * 1e-3) if r["ms"] > 0 else 0.0
    print(f"{name:22s} {r['calls']:5d} {r['ms']:9.3f} {r['bytes'] / 1e9:10.3f} {gbs:8.0f} {tf:8.1f}")
print(f"step: wall {wall_ms:.3f} ms, sum of bracketed calls {dev_ms:.3f} ms ({len(ops.records)} calls)")
