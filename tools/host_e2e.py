"""Time the host-buffer entry (am_host_gemm_strided_f32) on pinned buffers; AM_HOST_DEBUG=1 prints its timeline."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arraymancer_b200 import _capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
hA = torch.rand((n, n), dtype=torch.float32).pin_memory()
hB = torch.rand((n, n), dtype=torch.float32).pin_memory()
hC = torch.empty((n, n), dtype=torch.float32).pin_memory()
lib = _capi.lib()
for it in range(3):
    t0 = time.perf_counter()
    _capi.check(lib.am_host_gemm_strided_f32(n, n, n, 1.0, hA.data_ptr(), n, 1, hB.data_ptr(), n, 1, 0.0, hC.data_ptr(), n, 1))
    print(f"call {it}: {1e3 * (time.perf_counter() - t0):.1f} ms", file=sys.stderr)
