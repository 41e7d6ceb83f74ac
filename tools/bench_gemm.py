"""GPU tuning aid: times every GEMM shape of the train step under every tile shape of the tcgen05 kernel."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neraf_b200 import _lib  # noqa: E402

B = int(os.environ.get("BATCH", "2048"))
SHAPES = [  # (name, M, N, K, outputs)
    ("fwd1", B, 5096, 168, "rm+t"), ("fwd2", B, 2048, 5096, "rm+t"), ("fwd3", B, 1024, 2048, "rm+t"),
    ("fwd4", B, 1024, 1024, "rm+t"), ("fwd5", B, 512, 1024, "rm+t"), ("head", B, 513, 512, "f32"),
    ("dgrad_h", B, 512, 520, "rm+t"), ("dgrad5", B, 1024, 512, "rm+t"), ("dgrad4", B, 1024, 1024, "rm+t"),
    ("dgrad3", B, 2048, 1024, "rm+t"), ("dgrad2", B, 5096, 2048, "t"),
    ("wgrad_h", 513, 512, B, "f32"), ("wgrad5", 512, 1024, B, "f32"), ("wgrad4", 1024, 1024, B, "f32"),
    ("wgrad3", 1024, 2048, B, "f32"), ("wgrad2", 2048, 5096, B, "f32"), ("wgrad1", 5096, 163, B, "f32"),
]
TILES = [1064, 1128, 1256, 2064, 2128, 2256, 0]


def main():
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    print(f"{'shape':9s} {'M':>5s} {'N':>5s} {'K':>5s} " + " ".join(f"{t:>8d}" for t in TILES) + "   (us, best-of-5 after L2 flush)")
    total = {t: 0.0 for t in TILES}
    for name, M, N, K, outs in SHAPES:
        ldk = (K + 7) // 8 * 8
        A = torch.randn(M, ldk, device=dev).to(torch.bfloat16)
        Bm = torch.randn(N, ldk, device=dev).to(torch.bfloat16)
        e = _lib.GemmEpilogue()
        keep = []
        if "rm" in outs:
            o = torch.empty(M, (N + 7) // 8 * 8, dtype=torch.bfloat16, device=dev); keep.append(o)
            e.out_bf16, e.ld_bf16 = o.data_ptr(), o.stride(0)
        if "t" in outs:
            o = torch.empty(N, (M + 63) // 64 * 64, dtype=torch.bfloat16, device=dev); keep.append(o)
            e.out_bf16_t, e.ld_t = o.data_ptr(), o.stride(0)
        if "f32" in outs:
            o = torch.empty(M, N, device=dev); keep.append(o)
            e.out_f32, e.ld_f32 = o.data_ptr(), N
        row = []
        for t in TILES:
            _lib.check(lib.neraf_gemm_bf16_set_tile(t))
            best = 1e9
            for it in range(6):
                flush.fill_(0)
                s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                _lib.check(lib.neraf_gemm_bf16(M, N, K, A.data_ptr(), ldk, Bm.data_ptr(), ldk, C.byref(e), _lib.stream_ptr(dev)))
                f.record()
                torch.cuda.synchronize()
                if it:
                    best = min(best, s.elapsed_time(f) * 1e3)
            row.append(best)
            total[t] += best
        gf = 2.0 * M * N * K / 1e9
        print(f"{name:9s} {M:5d} {N:5d} {K:5d} " + " ".join(f"{v:8.1f}" for v in row) + f"   {gf:6.1f} GFLOP  best {gf / min(row) * 1e-3:6.0f} TFLOP/s")
    _lib.check(lib.neraf_gemm_bf16_set_tile(0))
    print("sum us  " + " " * 18 + " ".join(f"{total[t]:8.1f}" for t in TILES))


if __name__ == "__main__":
    main()
