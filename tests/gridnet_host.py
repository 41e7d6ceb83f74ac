"""Test infrastructure: ``GridOps`` on the CPU.  The neraf_grid_* operators are served by build/libgridnet_host.so -- a host
build of the very per-element code the CUDA kernels execute (neraf_b200/csrc/gridnet_core.h via tests/csrc/
gridnet_host.cpp) -- and the three GEMM forms by torch.matmul with fp32 accumulation and the output rounded to the
output tensor's dtype, as the tensor-core kernel does.  Lets tests/test_gridnet.py run neraf_b200/gridnet.py's assembly of
the ResNet3D against the reference network without a GPU.  Never imported by the product."""
import ctypes as C
import os


from neraf_b200 import _lib
from neraf_b200.gridnet import GridOps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_LIB = os.path.join(ROOT, "build", "libgridnet_host.so")


class HostOps(GridOps):
    prefix = "gridhost_"
    batched_backward = False           # gemm_backward -> gemm_tn + gemm_nn below

    def __init__(self):
        self.lib = C.CDLL(HOST_LIB)
        for name, (res, args) in _lib.GRID_SIGNATURES.items():
            fn = getattr(self.lib, name.replace("neraf_grid_", self.prefix))
            fn.restype, fn.argtypes = res, args
        self._counters = {}
        self.calls = {}

    def check_tensor(self, t, what):
        assert not t.is_cuda

    def stream(self, t):
        return None

    def call(self, name, *args):
        self.calls[name] = self.calls.get(name, 0) + 1
        assert getattr(self.lib, self.prefix + name)(*args) == 0

    def gemm_nt(self, A, B, M, N, K, out):
        self.calls["gemm_nt"] = self.calls.get("gemm_nt", 0) + 1
        out[:M, :N] = (A[:M, :K].float() @ B[:N, :K].float().t()).to(out.dtype)

    def gemm_nn(self, A, B, M, N, K, out):
        self.calls["gemm_nn"] = self.calls.get("gemm_nn", 0) + 1
        out[:M, :N] = (A[:M, :K].float() @ B[:K, :N].float()).to(out.dtype)

    def gemm_tn(self, A, B, M, N, K, out):
        self.calls["gemm_tn"] = self.calls.get("gemm_tn", 0) + 1
        out[:M, :N] = (A[:K, :M].float().t() @ B[:K, :N].float()).to(out.dtype)
