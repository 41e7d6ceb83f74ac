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
    batched_backward = True            # the product's job lists, split-K included, evaluated by run_gemms below

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

    def run_gemms(self, descs):
        for (M, N, K, A, B, a_mn, b_mn, out) in descs:
            kind = {(0, 0): "gemm_nt", (0, 1): "gemm_nn", (1, 1): "gemm_tn"}[(a_mn, b_mn)]
            self.calls[kind] = self.calls.get(kind, 0) + 1
            a = A[:K, :M].float().t() if a_mn else A[:M, :K].float()
            b = B[:K, :N].float() if b_mn else B[:N, :K].float().t()
            out[:M, :N] = (a @ b).to(out.dtype)
