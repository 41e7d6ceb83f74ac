"""GPU tuning aid: ONE large GEMM through the job-list kernel against torch.matmul (cuBLAS) on the same operands -- how far
the kernel's main loop (TMA ring -> tcgen05 -> epilogue) is from the library's when scheduling plays no role.

    python tools/gemm_peak.py [M N K]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neraf_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.lib()
counters = torch.zeros(1 << 16, dtype=torch.int32, device=dev)


def time_us(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    f.record()
    torch.cuda.synchronize()
    return s.elapsed_time(f) * 1e3 / reps


def main():
    shapes = [(8192, 8192, 8192), (4096, 4096, 4096), (16384, 2048, 5096), (16384, 5096, 2048), (2048, 2048, 5096)]
    if len(sys.argv) == 4:
        shapes = [tuple(int(a) for a in sys.argv[1:4])]
    for M, N, K in shapes:
        ldk = (K + 7) // 8 * 8
        A = (torch.randn(M, ldk, device=dev) * 0.05).to(torch.bfloat16)
        Bm = (torch.randn(N, ldk, device=dev) * 0.05).to(torch.bfloat16)
        out = torch.empty(M, (N + 7) // 8 * 8, dtype=torch.bfloat16, device=dev)
        row = []
        for bn in (256, 128):
            j = _lib.GemmJob()
            j.M, j.N, j.K, j.A, j.lda, j.B, j.ldb = M, N, K, A.data_ptr(), A.stride(0), Bm.data_ptr(), Bm.stride(0)
            j.bn, j.wait_job = bn, -1
            j.epi.out_bf16, j.epi.ld_bf16 = out.data_ptr(), out.stride(0)
            arr = (_lib.GemmJob * 1)(j)
            us = time_us(lambda: _lib.check(lib.neraf_gemm_bf16_jobs(arr, 1, counters.data_ptr(), counters.numel() * 4,
                                                                       _lib.stream_ptr(dev))))
            row.append((bn, us))
        At, Bt = A[:, :K], Bm[:, :K]
        us_t = time_us(lambda: torch.matmul(At, Bt.t()))
        gf = 2.0 * M * N * K / 1e6
        print(f"M={M} N={N} K={K}: " + "  ".join(f"bn={bn}: {us:8.1f} us {gf / us:7.0f} TFLOP/s" for bn, us in row)
              + f"   torch.matmul {us_t:8.1f} us {gf / us_t:7.0f} TFLOP/s", flush=True)
        ref = torch.matmul(At[:256].float(), Bt[:512].float().t())
        err = float((out[:256, :512].float() - ref).norm() / ref.norm())
        assert err < 1e-2, err


if __name__ == "__main__":
    main()
