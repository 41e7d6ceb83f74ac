"""GPU tuning aid: times the job-list kernel (neraf_gemm_bf16_jobs) layer by layer and as whole forward / backward lists.

    python tools/bench_jobs.py            # BATCH=2048 by default
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neraf_b200 import _lib  # noqa: E402

B = int(os.environ.get("BATCH", "2048"))
dev = torch.device("cuda:0")
lib = _lib.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
counters = torch.empty(1 << 16, dtype=torch.int32, device=dev)


def r8(x):
    return (x + 7) // 8 * 8


def bf(rows, cols):
    return (torch.randn(rows, r8(cols), device=dev) * 0.05).to(torch.bfloat16)


def job(M, N, K, A, Bm, bn, a_mn=0, b_mn=0, wait_job=-1, wait_all=0):
    j = _lib.GemmJob()
    j.M, j.N, j.K, j.A, j.lda, j.B, j.ldb = M, N, K, A.data_ptr(), A.stride(0), Bm.data_ptr(), Bm.stride(0)
    j.a_mn, j.b_mn, j.bn, j.wait_job, j.wait_all = a_mn, b_mn, bn, wait_job, wait_all
    return j


def run(jobs):
    arr = (_lib.GemmJob * len(jobs))(*jobs)
    _lib.check(lib.neraf_gemm_bf16_jobs(arr, len(jobs), counters.data_ptr(), counters.numel() * 4, _lib.stream_ptr(dev)))


def timeit(jobs, cold, reps=8):
    best = 1e9
    for it in range(reps):
        if cold:
            flush.fill_(0)
        s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        run(jobs)
        f.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, s.elapsed_time(f) * 1e3)
    return best


def choose_bn(M, N):
    rb = (M + 255) // 256
    for bn in (256, 128):
        if rb * ((N + bn - 1) // bn) >= 48:
            return bn
    return 64


def main():
    widths = [163, 5096, 2048, 1024, 1024, 512]
    CF = 513
    x = [bf(B, w) for w in widths]                  # activations x[0] = enc
    dz = [None] + [bf(B, w) for w in widths[1:]]
    w = [None] + [bf(widths[i + 1], widths[i]) for i in range(5)]      # (n, k)
    wt = [None] + [bf(widths[i], widths[i + 1]) for i in range(5)]     # (k, n)
    wh, wht = bf(CF, 512), bf(512, CF)
    dzh = bf(B, CF)
    out = torch.empty(B, CF, device=dev)
    dw = [None] + [torch.empty(widths[i + 1], widths[i], device=dev) for i in range(5)]
    dwh = torch.empty(CF, 512, device=dev)
    bias = torch.zeros(8192, device=dev)
    colsum = torch.zeros(8192, device=dev)

    def fwd_job(i, bn=None, wait=-1):
        if i < 5:
            j = job(B, widths[i + 1], widths[i], x[i], w[i + 1], bn or choose_bn(B, widths[i + 1]), wait_job=wait)
            j.epi.bias, j.epi.act = bias.data_ptr(), 1
            j.epi.out_bf16, j.epi.ld_bf16 = x[i + 1].data_ptr(), x[i + 1].stride(0)
        else:
            j = job(B, CF, 512, x[5], wh, bn or choose_bn(B, CF), wait_job=wait)
            j.epi.bias, j.epi.act = bias.data_ptr(), 2
            j.epi.out_f32, j.epi.ld_f32 = out.data_ptr(), CF
        return j

    def dgrad_job(i, bn=None, wait=-1):
        # dZ_{i} (B, widths[i]) = dZ_{i+1} W_{i+1}, i = 5 (from head) .. 1
        if i == 5:
            j = job(B, 512, CF, dzh, wht, bn or choose_bn(B, 512), wait_job=wait)
        else:
            j = job(B, widths[i], widths[i + 1], dz[i + 1], wt[i + 1], bn or choose_bn(B, widths[i]), wait_job=wait)
        j.epi.gate, j.epi.ldg = x[i].data_ptr(), x[i].stride(0)
        j.epi.out_bf16, j.epi.ld_bf16 = dz[i].data_ptr(), dz[i].stride(0)
        j.colsum = colsum.data_ptr()
        return j

    def wgrad_job(i, bn=None, wait=-1, wait_all=1):
        # dW_i (widths[i], widths[i-1]) = dZ_i^T x_{i-1}; i = 6: heads
        if i == 6:
            j = job(CF, 512, B, dzh, x[5], bn or max(128, choose_bn(CF, 512)), 1, 1, wait, wait_all)
            j.epi.out_f32, j.epi.ld_f32 = dwh.data_ptr(), 512
        else:
            j = job(widths[i], widths[i - 1], B, dz[i], x[i - 1], bn or max(128, choose_bn(widths[i], widths[i - 1])), 1, 1, wait, wait_all)
            j.epi.out_f32, j.epi.ld_f32 = dw[i].data_ptr(), dw[i].stride(0)
        return j

    if os.environ.get("NCU_SEQ"):
        # fixed launch sequence for `ncu --metrics gpu__time_duration.sum,...`: each entry launched 3 times
        t_a, t_b = bf(256, 64), bf(256, 64)
        t_o = torch.empty(256, 256, dtype=torch.bfloat16, device=dev)
        def tiny(n, k):
            a, b = bf(256, k), bf(n, k)
            j = job(256, n, k, a, b, 64 if n < 128 else (128 if n < 256 else 256))
            j.epi.out_bf16, j.epi.ld_bf16 = t_o.data_ptr(), 256
            return j, (a, b)
        seq = [("tiny256x64x64", [tiny(64, 64)[0]]), ("tiny256x256x64", [tiny(256, 64)[0]]), ("tiny256x256x4096", [tiny(256, 4096)[0]]),
               ("fwd6", [fwd_job(5)]), ("fwd1", [fwd_job(0)]), ("fwd2", [fwd_job(1)]), ("fwd3", [fwd_job(2)]), ("fwd4", [fwd_job(3)]),
               ("fwd5", [fwd_job(4)]), ("dgrad1", [dgrad_job(1)]), ("wgrad3", [wgrad_job(3, wait=-1, wait_all=0)]),
               ("wgrad2", [wgrad_job(2, wait=-1, wait_all=0)]),
               ("fwd_all", [fwd_job(i, None, i - 1) for i in range(6)])]
        keep = []
        for name, jobs in seq:
            for _ in range(3):
                run(jobs)
            torch.cuda.synchronize()
            print(name)
        return
    print(f"BATCH {B}: single-job launches (us; warm L2 / flushed L2), GFLOP, TFLOP/s warm")
    def report(name, mk, M, N, K, bns):
        gf = 2.0 * M * N * K / 1e9
        cells = []
        for bn in bns:
            tw, tc = timeit([mk(bn)], False), timeit([mk(bn)], True)
            cells.append(f"bn{bn}: {tw:7.1f}/{tc:7.1f} ({gf / tw * 1e3:5.0f} TF)")
        print(f"{name:8s} M{M:5d} N{N:5d} K{K:5d} {gf:6.1f} GF  " + "  ".join(cells))

    for i in range(6):
        N = widths[i + 1] if i < 5 else CF
        report(f"fwd{i + 1}", lambda bn, i=i: fwd_job(i, bn), B, N, widths[i], (64, 128, 256))
    for i in (5, 4, 3, 2, 1):
        K = CF if i == 5 else widths[i + 1]
        report(f"dgrad{i}", lambda bn, i=i: dgrad_job(i, bn), B, widths[i], K, (64, 128, 256))
    for i in (6, 5, 4, 3, 2, 1):
        M = CF if i == 6 else widths[i]
        N = 512 if i == 6 else widths[i - 1]
        report(f"wgrad{i}", lambda bn, i=i: wgrad_job(i, bn), M, N, B, (128, 256))

    fwd = [fwd_job(i, None, i - 1) for i in range(6)]
    print(f"forward list  (6 jobs)          warm {timeit(fwd, False):7.1f} us   cold {timeit(fwd, True):7.1f} us")
    print(f"forward 2..6  (5 jobs)          warm {timeit([fwd_job(i, None, i - 2) for i in range(1, 6)], False):7.1f} us")
    bwd = []
    prod = 0
    bwd.append(dgrad_job(5))
    bwd.append(wgrad_job(6, wait=-1, wait_all=0))
    for i in (5, 4, 3, 2, 1):
        dzp = prod
        if i > 1:
            prod = len(bwd)
            bwd.append(dgrad_job(i - 1, None, dzp))
        bwd.append(wgrad_job(i, None, dzp, 1))
    print(f"backward list ({len(bwd)} jobs)         warm {timeit(bwd, False):7.1f} us   cold {timeit(bwd, True):7.1f} us")
    chain = [j for j in bwd if not j.a_mn]
    # re-index the dgrad-only chain
    for k, j in enumerate(chain):
        j.wait_job = k - 1
    print(f"dgrad chain only ({len(chain)} jobs)      warm {timeit(chain, False):7.1f} us")
    wg = [wgrad_job(i, None, -1, 0) for i in (6, 5, 4, 3, 2, 1)]
    print(f"wgrad only ({len(wg)} jobs, no deps)  warm {timeit(wg, False):7.1f} us")
    tot_f = sum(2.0 * B * a * b for a, b in zip(widths[:-1], widths[1:])) + 2.0 * B * 512 * CF
    print(f"fwd GFLOP {tot_f / 1e9:.1f}  bwd GFLOP {(2 * tot_f - 2.0 * B * 163 * 5096) / 1e9:.1f}")


if __name__ == "__main__":
    main()
