#!/usr/bin/env python
"""Benchmark of the acoustic-field hot path (BASELINE.json metric: STFT columns/s train fwd+bwd, RIRs/s).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA library)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU torch path (oracle port)

One "step" = one pass of the hot path over one synthetic RAF-FurnishedRoom-shaped batch of 2048 STFT
columns per GPU: query encodings -> acoustic MLP -> spectral loss -> backward (+ gradient all-reduce for
N > 1).  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

FLOP_PER_COLUMN_TRAIN = {1: 2 * 44769136, 2: 2 * 44770672}       # SURVEY.md section 8d (factored layer 1)
GL_STREAM_BYTES_PER_RIR = {"RAF": 39.75e6, "SoundSpaces": 66.42e6}  # SURVEY.md section 8d streaming model
GL_COMPULSORY_BYTES_PER_RIR = {"RAF": 183536, "SoundSpaces": 306976}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU with nvidia-smi while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own algorithm (oracle port, torch CPU fp32 + autograd) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_train_step_factory(shape, batch_size, seed=0):
    from neraf_b200 import synthetic as syn
    from oracle import encodings as oenc, field as ofield, loss as oloss
    sd = {k: v.clone().requires_grad_(True) for k, v in syn.make_state_dict(shape, seed=seed).items()}
    batch = syn.make_batch(shape, batch_size, seed=seed)
    g = syn.make_grid_feature(seed).requires_grad_(True)
    aabb = syn.default_aabb()

    def step():
        for t in sd.values():
            t.grad = None
        g.grad = None
        h = oenc.assemble_input(batch, aabb, shape.T, g)                # NeRAF_model.py:533-560 (dense, like the reference)
        y = ofield.field_forward(sd, h, torch.float32)                   # NeRAF_field.py:47-65
        ld = oloss.loss_dict(y, batch["data"], "SC+SLMSE", 1e-3, dtype=torch.float32)
        (ld["audio_sc_loss"] + ld["audio_mag_loss"]).backward()
        return float(ld["audio_mag_loss"].detach())
    return step


def time_cpu_baseline(shape, batch_size, steps, warmup, budget_s=25.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_train_step_factory(shape, batch_size)
    for _ in range(warmup):
        step()
    times = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    sec = statistics.median(times)
    return {"value": batch_size / sec, "unit": "columns/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} train steps (fwd+loss+bwd, dense 1187-wide layer 1 as the reference executes it) of "
                      f"{batch_size} {shape.name}-shaped columns, torch {torch.__version__} CPU fp32, median",
            "ms_per_step": sec * 1e3, "steps": len(times)}


def run_reference(args):
    from neraf_b200 import synthetic as syn
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape = syn.RAF if args.shape == "RAF" else syn.SOUNDSPACES
    res = time_cpu_baseline(shape, args.batch, max(args.steps, 1), max(args.warmup, 1), budget_s=120.0)
    cfg_ref = workload_config(shape, args)
    cfg_ref.update({"precision": "fp32", "launch": "torch CPU eager, all host threads", "l2": "n/a (CPU)",
                    "grad_allreduce": "none (rank 0 only)",
                    "reference_kind": "port: /root/reference (and nerfstudio / tiny-cuda-nn, which its classes import) does not "
                                      "exist on the GPU box, so the arm times oracle/'s restatement of NeRAF_field.py / "
                                      "NeRAF_evaluator.py / the encodings -- pinned against the reference's own classes in the "
                                      "build container (oracle/make_golden.py)"})
    line = {"impl": "reference", "metric": "stft_columns_per_sec_train_fwd_bwd", "value": res["value"],
            "unit": "columns/s", "n_gpus": args.gpus, "steps": res["steps"], "warmup": max(args.warmup, 1),
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg_ref,
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(shape, args):
    return {"workload": f"{shape.name} FurnishedRoom-shaped audio-field training step (encode + MLP 1187->5096->2048->"
                        f"1024->1024->512->{shape.C}x{shape.F} + SC/log-STFT loss + backward), B={args.batch} columns/GPU, "
                        f"T={shape.T}", "batch_per_gpu": args.batch, "C": shape.C, "F": shape.F, "T": shape.T,
            "precision": args.precision,
            "launch": ("eager" if getattr(args, "no_graph", False) else
                       ("one CUDA graph per step" if (args.gpus == 1 or getattr(args, "kernel_exchange", False))
                        else "two CUDA graphs per step + eager NCCL all-reduces")) + " (value, e2e); eager plugin calls (e2e_eager)",
            "l2": "flushed between timed steps (256 MiB write, outside the events)",
            "grad_allreduce": ("none (one GPU)" if args.gpus == 1 else
                               ("bf16, by the library's own kernels over symmetric memory: the loss's four sums inside the loss "
                                "kernel, the gradients by a two-shot NVLS all-reduce (multimem.ld_reduce / multimem.st) running "
                                "beside the backward's GEMM launch -- no host-issued collective in the step"
                                if getattr(args, "kernel_exchange", False) else
                                f"{getattr(args, 'grad_dtype', 'fp32')} (one flat NCCL all-reduce after the backward)")),
            "optimizer": "not in the timed region (metric is fwd+bwd; nerfstudio's Adam is outside the path)",
            "grid_feature": "the batch-invariant 1024-vector is a learnable constant in `value` / `e2e` (its gradient dg is "
                            "computed); the ResNet3D-50 that produces it in the reference is timed in grid_feature.* and, inside "
                            "the train step, in step_with_producer",
            "tiles": "job-list kernel: static stride at this batch; explicit tile plans (csrc/mega_plan.h) where the planner's "
                     "model predicts >= 5 % (large_batch, batch_sweep)"}


# ------------------------------------------------------------------------------------------------
def arm_watchdog(seconds: float):
    """Hard exit if the run wedges (a hung collective must not hold the GPU box until the driver's limit)."""
    def fire():
        sys.stderr.write(f"bench.py watchdog: no result after {seconds:.0f} s, exiting\n")
        sys.stderr.flush()
        os._exit(3)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()


def main():
    arm_watchdog(900.0)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=2048, help="STFT columns per GPU per step (NeRAF_config.py:47)")
    ap.add_argument("--shape", default="RAF", choices=["RAF", "SoundSpaces"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--gl-rirs", type=int, default=2072, help="RIRs per Griffin-Lim launch (0 disables); 2072 = 14 per SM")
    ap.add_argument("--large-batch", type=int, default=16384, help="extra large-batch point of the sweep (0 disables)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--loss-columns", type=int, default=65536,
                    help="columns of the stand-alone spectral-loss timing (K2's HBM roofline: at the training batch the "
                         "kernels are launch-latency sized); 0 disables")
    ap.add_argument("--grid-net", type=int, default=128,
                    help="also time the grid-feature producer (ResNet3D-50 training-mode fwd+bwd, SURVEY 8f row 1) on an "
                         "N^3 grid (the reference's grid is 128^3); 0 disables")
    ap.add_argument("--no-graph", action="store_true", help="time the eager launch sequence instead of the CUDA graph")
    ap.add_argument("--exchange", default="auto", choices=["auto", "kernel", "nccl"],
                    help="N > 1: gradient exchange of the graphed step (GraphedTrainStep: the library's own kernels over "
                         "symmetric memory, or NCCL)")
    ap.add_argument("--sweep", default="1024,4096,8192,16384,32768,65536",
                    help="global batch sizes of the large-batch sweep (BASELINE config 5), split over the GPUs; '' disables")
    ap.add_argument("--no-soundspaces", action="store_true", help="skip the SoundSpaces-shaped training step (BASELINE config 3)")
    ap.add_argument("--grad-dtype", default=None, choices=["fp32", "bf16"],
                    help="dtype of the gradient all-reduce for N > 1 (default: bf16 with --precision bf16, else fp32)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.grad_dtype is None:
        args.grad_dtype = "bf16" if args.precision == "bf16" else "fp32"

    if args.impl == "reference":
        run_reference(args)
        return

    from neraf_b200 import _lib
    from neraf_b200 import synthetic as syn
    from neraf_b200.distributed import GradientAllReduce
    from neraf_b200.model import ConstantGridFeature, GraphedTrainStep, NeRAFAudioModel, NeRAFAudioModelConfig
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: neraf_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD if world > 1 else None

    shape = syn.RAF if args.shape == "RAF" else syn.SOUNDSPACES
    cfg = NeRAFAudioModelConfig(dataset=shape.name, max_len=shape.T, fs=shape.fs, N_freq_stft=shape.F,
                                hop_len=shape.hop, win_len=shape.win, precision=args.precision)
    model = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)),
                            process_group=group)
    model.field.load_state_dict(syn.make_state_dict(shape, seed=0))
    model = model.to(dev)
    model.field.always_repack = True          # training semantics: parameters change every step -> bf16 re-pack in every step
    params = [p for p in model.parameters() if p.requires_grad and p.numel() > 0]
    reducer = GradientAllReduce(params, group)

    B = args.batch
    host_batch = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, B, seed=rank).items()}
    dev_batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host_batch.items()}

    def step(batch):
        for p in params:
            p.grad = None
        out = model.get_outputs(batch)
        ld = model.get_loss_dict(out, batch)
        loss = sum(ld.values())
        loss.backward()
        reducer()
        return loss

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    lib = _lib.lib()

    # `value`: the same step captured in one CUDA graph (GraphedTrainStep), inputs resident in HBM.
    # `e2e`  : the plugin calls a nerfstudio Trainer makes, eager, host batch in pinned memory.
    graphed = None
    if not args.no_graph:
        # N = 1: one graph around the autograd calls.  N > 1: two graphs (forward + loss sums | backward) with the
        # loss's 32-byte all-reduce between them issued eagerly -- no NCCL call is ever captured
        graphed = GraphedTrainStep(model, dev_batch,
                                   grad_dtype=torch.bfloat16 if args.grad_dtype == "bf16" else torch.float32,
                                   exchange=args.exchange)
        args.kernel_exchange = graphed.kernel_exchange

    def step_value(batch):
        if graphed is None:
            return step(batch)
        ld = graphed(batch)
        graphed.allreduce_grads()
        return ld

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batch, n, read_loss, fn=None):
        """n steps, each bracketed by CUDA events on the launching stream, L2 flushed in between."""
        fn = fn or step
        evs = []
        barrier()
        l0 = lib.neraf_launch_count()
        t0 = time.perf_counter()
        for _ in range(n):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            loss = fn(batch)
            if read_loss:
                loss.item()                              # device -> host read of the step's result
            e.record()
            evs.append((s, e))
        barrier()
        wall = time.perf_counter() - t0
        ms = [s.elapsed_time(e) for s, e in evs]
        return ms, lib.neraf_launch_count() - l0, wall

    for _ in range(args.warmup):
        step_value(dev_batch)
        step(host_batch)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l_eager0 = lib.neraf_launch_count()
    step(dev_batch)
    launches_per_step = lib.neraf_launch_count() - l_eager0
    if graphed is not None:          # the graph replays the kernels of its last eager warm-up step
        launches_per_step = graphed.launches_per_step
    # `value`: inputs resident in HBM when the timed region starts -- the graphed step reads its static buffers in place
    ms_dev, launches, wall_dev = timed(graphed.static if graphed is not None else dev_batch, args.steps, read_loss=False,
                                       fn=step_value)
    if graphed is not None:
        launches = launches_per_step * args.steps
    ms_e2e_eager, _, wall_e2e = timed(host_batch, args.steps, read_loss=True)
    ms_e2e_g = ms_e2e_p = None
    if graphed is not None:          # same host batch through the repo's graphed step (H2D into static buffers + replay)
        def step_graphed_host(batch):
            step_value(batch)
            return graphed.total_loss
        ms_e2e_g, _, _ = timed(host_batch, args.steps, read_loss=True, fn=step_graphed_host)

        # ... and with the data loader's prefetch: every step consumes the batch staged by the previous step and
        # starts the host -> device copy of the next one (one copy per step, inside the timed region, on a copy stream)
        def step_graphed_prefetch(batch):
            step_value(batch)
            graphed.prefetch(batch)
            return graphed.total_loss
        graphed.prefetch(host_batch)
        for _ in range(3):
            step_graphed_prefetch(host_batch)
        ms_e2e_p, _, _ = timed(host_batch, args.steps, read_loss=True, fn=step_graphed_prefetch)
    ms_e2e = ms_e2e_p if ms_e2e_p is not None else ms_e2e_eager
    ms_e2e_feed = None
    if graphed is not None:
        # ... and with the GPU-resident data feed (neraf_b200.datafeed): the batch is gathered on the device from a
        # cache of target columns straight into the graph's static buffers; nothing crosses PCIe but the loss
        from neraf_b200.datafeed import ResidentAudioFeed
        n_cache = 2048
        gf = torch.Generator().manual_seed(77 + rank)
        cache = (torch.randn(n_cache, shape.T, shape.C, shape.F, generator=gf) - 3.0).to(dev)
        pose = lambda: syn.make_batch(shape, n_cache, seed=5 + rank)          # noqa: E731
        pb = pose()
        feed = ResidentAudioFeed(cache, pb["mic_pose"], pb["source_pose"], pb["rot"], shape.T, B, seed=1, rank=rank,
                                 world_size=world, drop_last=True)
        counter = [0]

        def step_graphed_feed(_):
            feed.next_train(counter[0], out=graphed.static)
            counter[0] += 1
            step_value(graphed.static)
            return graphed.total_loss
        for _ in range(3):
            step_graphed_feed(None)
        ms_e2e_feed, _, _ = timed(None, args.steps, read_loss=True, fn=step_graphed_feed)
        feed.check()
        del cache
    clocks = sampler.stop()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    total_ms = max_over_ranks(sum(ms_dev))
    total_ms_e2e = max_over_ranks(sum(ms_e2e))
    ms_per_step = total_ms / args.steps
    value = B * world * args.steps / (total_ms * 1e-3)
    e2e_value = B * world * args.steps / (total_ms_e2e * 1e-3)

    def e2e_entry(ms, api):
        tot = max_over_ranks(sum(ms))
        return {"value": B * world * args.steps / (tot * 1e-3), "unit": "columns/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": tot / args.steps, "api": api}

    e2e_api = ("GraphedTrainStep(pinned host batch) -> prefetch(next pinned host batch) -> loss.item(): one host -> device "
               "copy of a batch per step on a copy stream (overlapping the step's kernels), staged -> static buffers on "
               "the device, graph replay, 4-byte loss read") if ms_e2e_p is not None else \
              "NeRAFAudioModel.get_outputs(pinned host batch) -> get_loss_dict -> backward -> loss.item() (eager)"
    peaks = load_peaks()
    flops_step = FLOP_PER_COLUMN_TRAIN[shape.C] * B
    achieved = flops_step / (ms_per_step * 1e-3) / 1e12
    h2d = sum(v.numel() * v.element_size() for k, v in host_batch.items()
              if torch.is_tensor(v) and k in ("time_query", "mic_pose", "source_pose", "rot", "data"))

    line = {
        "metric": "stft_columns_per_sec_train_fwd_bwd", "value": value, "unit": "columns/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": workload_config(shape, args), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "columns/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": total_ms_e2e / args.steps,
                "api": e2e_api},
        "e2e_graphed_no_prefetch": None if ms_e2e_g is None else e2e_entry(
            ms_e2e_g, "GraphedTrainStep(pinned host batch) -> loss.item(): the batch is copied into the graph's static "
                      "buffers on the compute stream, then the graph replays"),
        "e2e_resident_feed": None if ms_e2e_feed is None else dict(e2e_entry(
            ms_e2e_feed, "ResidentAudioFeed.next_train(out=static buffers) -> GraphedTrainStep -> loss.item(): batches "
                         "gathered on the device from a resident cache of target columns (neraf_gather_batch), the "
                         "epoch permutation uploaded once per epoch"), h2d_bytes_per_step=8 * B),
        "e2e_eager": e2e_entry(
            ms_e2e_eager, "NeRAFAudioModel.get_outputs(pinned host batch) -> get_loss_dict -> backward -> loss.item(): the "
                          "plugin calls a nerfstudio Trainer makes, one C-ABI call per autograd node"),
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_tflops_sustained"],
                     # dram__bytes_read.sum + dram__bytes_write.sum of the step's two job-list launches at B=2048 RAF bf16
                     # (ncu --set full, profiles/r01e_mega_ncu_summary.txt: 49.9 MB forward + 172.5 MB backward)
                     "traffic": 209.2e6 if (B == 2048 and shape.name == "RAF" and args.precision == "bf16") else None,   # ncu, profiles/r02ap_mega_ncu_summary_variants.txt: 44.7 + 164.5 MB
                     "traffic_unit": "bytes per step (both umma_mega_kernel launches)",
                     "kernel": "umma_mega_kernel (job-list tcgen05 kernel: all GEMMs of the step in 3 launches; achieved = 89.54 "
                               "MFLOP/column x columns / whole-step device time, i.e. the non-GEMM kernels of the step are "
                               "charged to it)",
                     "peak_source": f"{peaks['source']} sustained bf16 cuBLAS (MEASURED_PEAKS.json)"},
        "host_wall_ms_per_step": wall_dev / args.steps * 1e3,
    }

    # ---- the same step at 8x the reference batch (BASELINE config 5's sweep, one point): what the kernels reach once
    # a layer is more than one tile per CTA pair.  Not the headline: `value` stays at the reference batch.
    if world == 1 and graphed is not None and args.large_batch > 0 and args.large_batch != B:
        BL = args.large_batch
        big = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, BL, seed=3).items()}
        g_big = GraphedTrainStep(model, big)
        for _ in range(3):
            g_big(big)
        ms_big, _, _ = timed(big, 20, read_loss=False, fn=g_big)
        ms_b = sum(ms_big) / len(ms_big)
        ach = FLOP_PER_COLUMN_TRAIN[shape.C] * BL / (ms_b * 1e-3) / 1e12
        line["large_batch"] = {"batch_per_gpu": BL, "value": BL / (ms_b * 1e-3), "unit": "columns/s", "ms_per_step": ms_b,
                               "roofline_frac": ach / peaks["bf16_tflops_sustained"], "achieved_tflops": ach, "steps": 20}
        del g_big, big

    def time_graphed(shape_x, model_x, b, steps_x, seed):
        """ms per step (max over ranks) of the graphed training step at b columns per GPU: inputs resident in the step's
        static buffers, L2 flushed between steps, CUDA events per step."""
        bt = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape_x, b, seed=seed + rank).items()}
        g = GraphedTrainStep(model_x, bt, exchange=args.exchange)

        def fn(_):
            g(g.static)
            g.allreduce_grads()
        for _ in range(3):
            fn(None)
        ms_x, _, _ = timed(None, steps_x, read_loss=False, fn=fn)
        tot = max_over_ranks(sum(ms_x))
        kx = g.kernel_exchange
        del g, bt
        return tot / steps_x, kx

    # ---- BASELINE config 5: the large-batch sweep, GLOBAL batch split over the GPUs (strong scaling per point)
    if graphed is not None and args.sweep:
        rows = []
        for gb in [int(x) for x in args.sweep.split(",") if x]:
            b = gb // world
            if b < 128:
                continue
            ms_b, kx = time_graphed(shape, model, b, 20, seed=31)
            ach = FLOP_PER_COLUMN_TRAIN[shape.C] * b / (ms_b * 1e-3) / 1e12
            rows.append({"global_batch": gb, "batch_per_gpu": b, "ms_per_step": ms_b, "value": b * world / (ms_b * 1e-3),
                         "unit": "columns/s", "roofline_frac_per_gpu": ach / peaks["bf16_tflops_sustained"]})
        line["batch_sweep"] = {"scaling": "strong (global batch / n_gpus columns per GPU)", "steps": 20, "points": rows,
                               "note": "same graphed step as `value`; inputs resident in the static buffers"}

    # ---- BASELINE config 3: SoundSpaces-shaped (binaural, orientation-conditioned) training step, data parallel over N
    if graphed is not None and not args.no_soundspaces and shape.name != "SoundSpaces":
        sx = syn.SOUNDSPACES
        cfg_x = NeRAFAudioModelConfig(dataset=sx.name, max_len=sx.T, fs=sx.fs, N_freq_stft=sx.F, hop_len=sx.hop,
                                      win_len=sx.win, precision=args.precision)
        model_x = NeRAFAudioModel(cfg_x, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)),
                                  process_group=group)
        model_x.field.load_state_dict(syn.make_state_dict(sx, seed=0))
        model_x = model_x.to(dev)
        model_x.field.always_repack = True
        ms_x, kx = time_graphed(sx, model_x, B, 50, seed=41)
        ach = FLOP_PER_COLUMN_TRAIN[sx.C] * B / (ms_x * 1e-3) / 1e12
        line["soundspaces"] = {"metric": "stft_columns_per_sec_train_fwd_bwd", "value": B * world / (ms_x * 1e-3),
                               "unit": "columns/s", "ms_per_step": ms_x, "batch_per_gpu": B, "C": sx.C, "F": sx.F, "T": sx.T,
                               "n_gpus": world, "scaling": "weak", "steps": 50,
                               "roofline_frac_per_gpu": ach / peaks["bf16_tflops_sustained"],
                               "exchange": "kernel" if kx else ("nccl" if world > 1 else "none")}
        del model_x

    # ---- N > 1: the data-parallel step must BE single-process training on the concatenated batch (the
    # reference refuses world_size > 1, so this equality is the specification); a mismatch fails the run
    if world > 1 and graphed is not None:
        Bc = 768
        keys = ("time_query", "mic_pose", "source_pose", "rot", "data")
        mine = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.make_batch(shape, Bc, seed=910 + rank).items()}
        whole = {}
        for k in keys:
            parts = [torch.empty_like(mine[k]) for _ in range(world)]
            dist.all_gather(parts, mine[k].contiguous())
            whole[k] = torch.cat(parts)
        single = NeRAFAudioModel(cfg, syn.default_aabb(), resnet3d=ConstantGridFeature(1024, syn.make_grid_feature(0)))
        single.field.load_state_dict(syn.make_state_dict(shape, seed=0))
        single = single.to(dev)
        ld_s = single.get_loss_dict(single.get_outputs(whole), whole)
        sum(ld_s.values()).backward()
        flat = lambda m: torch.cat([p.grad.reshape(-1).float() for p in m.parameters() if p.requires_grad and p.numel() > 0])  # noqa: E731
        ref_g = flat(single)
        g_chk = GraphedTrainStep(model, mine, exchange=args.exchange)
        for _ in range(3):
            got = g_chk(mine)
            g_chk.allreduce_grads()
        torch.cuda.synchronize()
        err = float((flat(model) - ref_g).norm() / ref_g.norm())
        lerr = max(abs(float(got[k]) - float(ld_s[k])) / abs(float(ld_s[k])) for k in got)
        tol = 4e-3 if args.precision == "bf16" else 1e-4          # bf16 exchange: 2^-9 per addend
        err, lerr = max_over_ranks(err), max_over_ranks(lerr)
        line["dp_equals_single"] = {"ranks": world, "columns_per_rank": Bc, "summed_gradients_rel_err": err, "losses_rel_err": lerr,
                                    "tolerance": tol, "exchange": "kernel" if g_chk.kernel_exchange else "nccl",
                                    "ok": bool(err < tol and lerr < 1e-5)}
        del g_chk, single
        if not line["dp_equals_single"]["ok"]:
            if rank == 0:
                sys.stderr.write(f"bench.py: data-parallel step != single-process step: {line['dp_equals_single']}\n")
            raise SystemExit(4)

    # ---- Griffin-Lim: RIRs/s (second half of the BASELINE metric), rank-local poses, no collective
    if args.gl_rirs > 0:
        n = args.gl_rirs
        gen = torch.Generator().manual_seed(100 + rank)
        log_h = (torch.randn(n, shape.T, shape.C, shape.F, generator=gen) * 1.5 - 3.0).pin_memory()
        log_d = log_h.to(dev)
        init = torch.rand(n, shape.C, shape.F, shape.T, dtype=torch.complex64, device=dev)
        gl = model.istft_transform
        for _ in range(2):
            gl.render(log_d, init)
        barrier()
        k_gl = 3
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(k_gl):
            gl.render(log_d, init)
        e.record()
        barrier()
        gl_ms = max_over_ranks(s.elapsed_time(e)) / k_gl
        s.record()
        for _ in range(k_gl):
            w = gl.render(log_h.to(dev, non_blocking=True), init)
            w_host = w.cpu()
        e.record()
        barrier()
        gl_ms_e2e = max_over_ranks(s.elapsed_time(e)) / k_gl
        rirs = n * world / (gl_ms * 1e-3)
        stream_gbs = rirs / world * GL_STREAM_BYTES_PER_RIR[shape.name] / 1e9
        line["griffinlim"] = {
            "metric": "rirs_per_sec", "value": rirs, "unit": "RIR/s", "rirs_per_launch_per_gpu": n, "ms_per_launch": gl_ms,
            "e2e": {"value": n * world / (gl_ms_e2e * 1e-3), "unit": "RIR/s",
                    "h2d_bytes_per_step": log_h.numel() * 4, "d2h_bytes_per_step": w_host.numel() * 4},
            "roofline": {"bound": "hbm", "achieved": stream_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": stream_gbs / peaks["hbm_gbs"], "traffic": None,
                         "note": "achieved = streaming-model algorithmic bytes (SURVEY.md 8d: 39.75 MB/RIR RAF) x RIR/s; the "
                                 "fused kernel keeps the state on chip, its compulsory HBM bytes are "
                                 f"{GL_COMPULSORY_BYTES_PER_RIR[shape.name]} B/RIR"}}

        # ---- batched RIR rendering (BASELINE config 4, loudness-map workload): poses -> field query over all T time
        # bins -> log->magnitude -> Griffin-Lim, one grid feature for all poses; rank-local poses, no collective
        n_pose = 1024
        gp = torch.Generator().manual_seed(200 + rank)
        aabb = syn.default_aabb()
        lo, hi = aabb[0] + 1.0, aabb[1] - 1.0
        mic = (lo + (hi - lo) * torch.rand(n_pose, 3, generator=gp)).double().pin_memory()
        src = ((lo + hi) / 2).double().reshape(1, 3)
        rot = torch.tensor([[1.0, 0.5, 0.5]], dtype=torch.float64)
        init_r = init[:n_pose] if n >= n_pose else torch.rand(n_pose, shape.C, shape.F, shape.T, dtype=torch.complex64, device=dev)
        model.field.always_repack = False
        wr = torch.empty(n_pose, shape.C, shape.hop * (shape.T - 1), dtype=torch.float32).pin_memory()
        for _ in range(2):
            wr.copy_(model.render_rirs(mic, src, rot, init_r), non_blocking=True)
        barrier()
        s.record()
        for _ in range(k_gl):
            wr.copy_(model.render_rirs(mic, src, rot, init_r), non_blocking=True)      # waveforms into pinned host memory
        e.record()
        barrier()
        r_ms = max_over_ranks(s.elapsed_time(e)) / k_gl
        model.field.always_repack = True
        line["render"] = {"metric": "rendered_rirs_per_sec", "value": n_pose * world / (r_ms * 1e-3), "unit": "RIR/s",
                          "poses_per_call_per_gpu": n_pose, "queries_per_call_per_gpu": n_pose * shape.T, "ms_per_call": r_ms,
                          "api": "NeRAFAudioModel.render_rirs(host poses) -> waveforms on the host (field forward over T bins "
                                 "+ Griffin-Lim)", "d2h_bytes_per_call": wr.numel() * 4}

        # ---- acoustic metrics of the rendered RIRs (T60 / EDT / C50), one launch for the whole batch
        from neraf_b200.metrics import acoustic_metrics
        wd = gl.render(log_d, init)                                    # (n, C, L) on the device
        advanced = shape.C == 1
        for _ in range(2):
            acoustic_metrics(wd, shape.fs, advanced)
        barrier()
        s.record()
        for _ in range(k_gl):
            mt = acoustic_metrics(wd, shape.fs, advanced)
            host_m = {k: v.cpu() for k, v in mt.items()}
        e.record()
        barrier()
        m_ms = max_over_ranks(s.elapsed_time(e)) / k_gl
        line["acoustic_metrics"] = {"metric": "rirs_measured_per_sec", "value": n * world / (m_ms * 1e-3), "unit": "RIR/s",
                                    "rirs_per_call_per_gpu": n, "ms_per_call": m_ms,
                                    "api": "neraf_b200.metrics.acoustic_metrics(device waveforms) -> T60, EDT, C50 on the host",
                                    "d2h_bytes_per_call": sum(v.numel() * 8 for v in host_m.values())}
        del wd

    # ---- spectral loss alone at a size where it is bandwidth- and not latency-sized ("measure at 64 k too"):
    # forward = sums + finalize in one launch (8 B / element read), backward = gradient (8 B read + 4 B written);
    # pred and gt (2 x 134 MB at 65 536 RAF columns) are larger than L2, so every launch streams from HBM
    if args.loss_columns > 0:
        sl, sl_err = (-1.0, -1.0), None
        try:
            n_el = args.loss_columns * shape.C * shape.F
            pred_l = torch.randn(n_el, device=dev) * 1.5 - 3.0
            gt_l = torch.randn(n_el, device=dev) * 1.5 - 3.0
            scratch_l = torch.zeros(5, dtype=torch.float64, device=dev)
            losses_l = torch.zeros(2, dtype=torch.float32, device=dev)
            dpred_l = torch.empty_like(pred_l)
            crit, st = _lib.CRITERIA["SC+SLMSE"], _lib.stream_ptr(dev)

            def loss_fwd():
                _lib.check(lib.neraf_spectral_loss_forward(pred_l.data_ptr(), gt_l.data_ptr(), n_el, crit, 1e-4, 1e-3,
                                                           scratch_l.data_ptr(), losses_l.data_ptr(), st))

            def loss_bwd():
                _lib.check(lib.neraf_spectral_loss_backward(pred_l.data_ptr(), gt_l.data_ptr(), n_el, n_el, crit,
                                                            scratch_l.data_ptr(), None, None, 1e-4, 1e-3,
                                                            dpred_l.data_ptr(), st))
            out_ms = []
            for fn in (loss_fwd, loss_bwd):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(20):
                    fn()
                e.record()
                torch.cuda.synchronize()
                out_ms.append(s.elapsed_time(e) / 20)
            sl = tuple(out_ms)
            del pred_l, gt_l, dpred_l
        except Exception as exc:                                   # noqa: BLE001 -- an extra row must not lose the line
            sl_err = f"{type(exc).__name__}: {exc}"
        sl_failed = max_over_ranks(1.0 if sl_err else 0.0) > 0
        f_ms, b_ms = max_over_ranks(sl[0]), max_over_ranks(sl[1])
        if sl_failed:
            line["spectral_loss"] = {"error": sl_err or "another rank failed"}
        else:
            n_el = args.loss_columns * shape.C * shape.F
            f_gbs, b_gbs = 8.0 * n_el / (f_ms * 1e-3) / 1e9, 12.0 * n_el / (b_ms * 1e-3) / 1e9
            line["spectral_loss"] = {
                "columns_per_gpu": args.loss_columns, "elements_per_gpu": n_el,
                "forward": {"ms": f_ms, "bytes_per_element": 8, "gbs": f_gbs, "frac": f_gbs / peaks["hbm_gbs"]},
                "backward": {"ms": b_ms, "bytes_per_element": 12, "gbs": b_gbs, "frac": b_gbs / peaks["hbm_gbs"]},
                "roofline": {"bound": "hbm", "achieved": f_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": f_gbs / peaks["hbm_gbs"],
                             # dram__bytes_read + write of one launch (ncu --set full, profiles/r02z_k2_ncu_summary.txt: 269.0 +
                             # 4.4 MB at 33.6 M elements = the 8 B/element of the algorithm, nothing is read twice)
                             "traffic": 273.4e6 if (args.loss_columns == 65536 and shape.name == "RAF") else None,
                             "traffic_backward": 362.4e6 if (args.loss_columns == 65536 and shape.name == "RAF") else None,
                             "note": "neraf_spectral_loss_forward (pred + gt read once, fp64 accumulation), 20 back-to-back "
                                     "launches on inputs larger than L2; per GPU; the backward's ncu capture still held 41 MB of "
                                     "its 134 MB of writes in L2 when it ended"}}

    # ---- grid-feature producer: one training-mode forward + backward of ResNet3D-50 on a (1, 7, N, N, N) grid
    if args.grid_net > 0:
        # (no collective inside the try block: a rank that fails must not leave the others waiting in one)
        gn_local, gn_err, gn_info = -1.0, None, {}
        try:
            from neraf_b200.gridnet import ResNet3D_helper, conv_flops
            n_g = args.grid_net
            net = ResNet3D_helper(in_channels=7, backbone="resnet50", grid_step=1.0 / n_g, N_features=1024,
                                  precision=args.precision)
            net.load_state_dict(syn.make_gridnet_state_dict("resnet50"))
            net = net.to(dev).train()
            grid = syn.make_grid(n_g).to(dev)
            dfeat = torch.randn(1, 1024, 1, 1, 1, device=dev)
            for _ in range(2):
                net(grid).backward(dfeat)
            torch.cuda.synchronize()
            l0 = lib.neraf_launch_count()
            k_gn = 5
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(k_gn):
                net(grid).backward(dfeat)
            e.record()
            torch.cuda.synchronize()
            gn_local = s.elapsed_time(e) / k_gn
            gn_info = {"gflop": conv_flops(net.backbone_net, n_g) / 1e9,
                       "launches": (lib.neraf_launch_count() - l0) // k_gn}
            del net, grid
        except Exception as exc:                                   # noqa: BLE001 -- an extra row must not lose the line
            gn_err = f"{type(exc).__name__}: {exc}"
        any_failed = max_over_ranks(1.0 if gn_err else 0.0) > 0
        gn_ms = max_over_ranks(gn_local)
        if any_failed:
            line["grid_feature"] = {"error": gn_err or "another rank failed"}
        else:
            line["grid_feature"] = {"metric": "grid_net_steps_per_sec", "value": world / (gn_ms * 1e-3), "unit": "step/s",
                                    "grid": [1, 7, args.grid_net] + [args.grid_net] * 2, "ms_per_step": gn_ms,
                                    "gflop_per_step": gn_info["gflop"],
                                    "achieved_tflops": gn_info["gflop"] / 1e3 / (gn_ms * 1e-3),
                                    "launches_per_step": gn_info["launches"],
                                    "api": "gridnet.ResNet3D_helper(grid).backward(): training-mode batch norm, all parameter gradients"}
        # the same step captured in ONE CUDA graph (no Python / ctypes between the ~440 launches), timed in a child
        # process: a capture has never run on this path before, and a failure there must not cost this process its line
        if world == 1 and "error" not in line["grid_feature"]:
            try:
                child = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools",
                                                                     "gridnet_quick.py"), str(args.grid_net), args.precision,
                                        "--graph"], capture_output=True, text=True, timeout=180)
                rows = [ln for ln in child.stdout.splitlines() if ln.startswith("{")]
                got = json.loads(rows[-1]) if rows else {}
                if "graph_ms_per_step" in got:
                    line["grid_feature"]["graphed"] = {
                        "ms_per_step": got["graph_ms_per_step"], "achieved_tflops": got["graph_achieved_tflops"],
                        "value": 1.0 / (got["graph_ms_per_step"] * 1e-3), "unit": "step/s",
                        "grad_max_rel_vs_eager": got.get("graph_vs_eager_grad_max_rel"),
                        "api": "torch.cuda.CUDAGraph over ResNet3D_helper(grid).backward() (tools/gridnet_quick.py --graph)"}
                else:
                    line["grid_feature"]["graphed"] = {"error": got.get("graph_error") or (child.stderr or "no output")[-300:]}
            except Exception as exc:                               # noqa: BLE001
                line["grid_feature"]["graphed"] = {"error": f"{type(exc).__name__}: {exc}"}
            # ... and the audio train step AS THE REFERENCE RUNS IT: producer forward -> field step -> producer backward in
            # one graph (NeRAF_model.py:554-566: the ResNet3D runs, and trains, in every audio step).  The headline `value`
            # holds the grid feature constant (BASELINE's metric is the acoustic-field path); this is the step with it.
            try:
                child = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools",
                                                                     "step_with_producer.py"), str(args.grid_net), str(B)],
                                       capture_output=True, text=True, timeout=240)
                rows = [ln for ln in child.stdout.splitlines() if ln.startswith("{")]
                line["step_with_producer"] = json.loads(rows[-1]) if rows else {"error": (child.stderr or "no output")[-300:]}
            except Exception as exc:                               # noqa: BLE001
                line["step_with_producer"] = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if args.gl_rirs > 0:
            from oracle import metrics as omet
            n_cpu = 32
            rir_c, _, _ = syn.make_rirs(shape, n_cpu, seed=0)
            hs = rir_c.reshape(-1, rir_c.shape[-1]).numpy().astype(np.float32)
            t0 = time.perf_counter()
            for h1 in hs:
                (omet.t60_raf if shape.C == 1 else omet.t60_soundspaces)(h1, shape.fs)
                omet.measure_edt(h1, shape.fs)
                omet.measure_clarity(h1, fs=shape.fs)
            dt = time.perf_counter() - t0
            line["acoustic_metrics"]["cpu_baseline"] = {
                "value": n_cpu / dt, "unit": "RIR/s", "cores": 1, "kind": "port",
                "sample": f"{n_cpu} {shape.name}-shaped RIRs, oracle restatement of NeRAF_helper.py compute_t60 / "
                          "measure_edt / measure_clarity (numpy, one RIR at a time like the reference)"}
        res = time_cpu_baseline(shape, B, steps=10, warmup=2)
        line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if args.grid_net > 0 and "error" not in line.get("grid_feature", {"error": 1}):
            from oracle import gridnet as ogn
            n_c = args.grid_net                        # the SAME grid as the GPU figure (about half a minute of CPU work at 128^3)
            sd_c, x_c = syn.make_gridnet_state_dict("resnet50"), syn.make_grid(n_c)
            t0 = time.perf_counter()
            ogn.forward_backward(sd_c, x_c, torch.ones(1, 1024, 1, 1, 1), 1.0 / n_c, dtype=torch.float32)
            dt = time.perf_counter() - t0
            line["grid_feature"]["cpu_baseline"] = {
                "value": 1.0 / dt, "unit": "step/s", "cores": os.cpu_count() or 1, "kind": "port", "grid": [1, 7, n_c, n_c, n_c],
                "sample": f"one training-mode fwd+bwd of the oracle (torch {torch.__version__} CPU fp32 conv3d / batch_norm) on a "
                          f"{n_c}^3 grid (the GPU figure's grid)"}
        if args.gl_rirs > 0:
            from oracle import griffinlim as ogl
            n_cpu = 16
            _, mag_c, init_c = syn.make_rirs(shape, n_cpu, seed=0)
            t0 = time.perf_counter()
            ogl.griffinlim(mag_c, init_c, shape.n_fft, shape.hop, shape.win)
            dt = time.perf_counter() - t0
            line["griffinlim"]["cpu_baseline"] = {"value": n_cpu / dt, "unit": "RIR/s", "cores": os.cpu_count() or 1, "kind": "port",
                                                  "sample": f"{n_cpu} {shape.name}-shaped RIRs, 32 iterations, oracle restatement of "
                                                            f"torchaudio GriffinLim on torch {torch.__version__} CPU"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
