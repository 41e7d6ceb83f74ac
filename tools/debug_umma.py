"""GPU debugging aid (not a test): runs the tcgen05 GEMM on small shapes and prints error structure."""
import ctypes as C
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neraf_b200 import _lib  # noqa: E402


def run(M, N, K, seed=0):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g).to(dev).to(torch.bfloat16)
    B = torch.randn(N, K, generator=g).to(dev).to(torch.bfloat16)
    out = torch.full((M, N), -777.0, device=dev)
    e = _lib.GemmEpilogue()
    e.out_f32, e.ld_f32 = out.data_ptr(), N
    rc = _lib.lib().neraf_gemm_bf16(M, N, K, A.data_ptr(), K, B.data_ptr(), K, C.byref(e), _lib.stream_ptr(dev))
    if rc:
        print(f"M={M} N={N} K={K}: rc={rc} {_lib.lib().neraf_last_error().decode()}")
        return
    try:
        torch.cuda.synchronize()
    except Exception as ex:  # noqa: BLE001
        print(f"M={M} N={N} K={K}: CUDA error {ex}")
        raise
    ref = A.double() @ B.double().t()
    err = (out.double() - ref).abs()
    rel = float(err.max() / ref.abs().max())
    untouched = int((out == -777.0).sum())
    print(f"M={M} N={N} K={K}: max rel err {rel:.3e} untouched {untouched}")
    if rel > 1e-3:
        bad_rows = (err.max(dim=1).values > 1e-2 * ref.abs().max()).nonzero().flatten().tolist()
        bad_cols = (err.max(dim=0).values > 1e-2 * ref.abs().max()).nonzero().flatten().tolist()
        print("  bad rows:", bad_rows[:16], "... n=", len(bad_rows), " bad cols:", bad_cols[:16], "... n=", len(bad_cols))
        print("  out[0,:8]", out[0, :8].tolist())
        print("  ref[0,:8]", ref[0, :8].tolist())
        # is it a K-subset problem? compare against partial sums over k-chunks of 16
        for kk in range(16, K + 1, 16):
            part = A[:, :kk].double() @ B[:, :kk].double().t()
            if float((out.double() - part).abs().max() / ref.abs().max()) < 1e-3:
                print(f"  output equals the partial sum over the first {kk} of {K} k-elements")
                break


if __name__ == "__main__":
    print("device", torch.cuda.get_device_name(0), "supported", _lib.lib().neraf_device_supported())
    if len(sys.argv) > 1:
        _lib.check(_lib.lib().neraf_gemm_bf16_set_tile(int(sys.argv[1])))
        print("tile override", sys.argv[1])
    for shape in [(128, 64, 64), (128, 128, 64), (128, 256, 64), (128, 64, 16), (128, 64, 128), (128, 64, 512),
                  (256, 256, 256), (2048, 2048, 512), (100, 72, 40), (2048, 5096, 163 + 5)]:
        run(*shape)
