"""Single-process multi-GPU entries of the C ABI (am_mg_*, include/am_b200.h): one host process drives the GPUs of a
box — the form a Nim caller uses (the reference never calls cudaSetDevice; these entries manage devices themselves and
restore the caller's).  The one-process-per-GPU form (torch.distributed) is in distributed.py.

Row-sharded GEMM of SURVEY 8e: GPU g owns the rows `rows(M, g)` of A and C, B is replicated, every GPU ends up with
the whole C.  torch tensors only carry device memory."""
from __future__ import annotations

import ctypes

import torch

from . import _capi
from .cuda_tensor import _SUFFIX


class MgContext:
    def __init__(self, devices=None, ndev: int | None = None):
        if devices is None:
            ndev = ndev if ndev is not None else torch.cuda.device_count()
            devices = list(range(ndev))
        self.devices = [int(d) for d in devices]
        arr = (ctypes.c_int * len(self.devices))(*self.devices)
        self._h = ctypes.c_void_p()
        _capi.check(_capi.lib().am_mg_init(len(self.devices), arr, ctypes.byref(self._h)))

    @property
    def ndev(self) -> int:
        return len(self.devices)

    def rows(self, M: int, g: int):
        r0, n = ctypes.c_int64(), ctypes.c_int64()
        _capi.check(_capi.lib().am_mg_rows(self._h, M, g, ctypes.byref(r0), ctypes.byref(n)))
        return r0.value, n.value

    def gemm_rowsharded(self, alpha, A_local, B, C):
        """A_local[g]: [rows_g, K] on GPU g (row-major, any row pitch); B[g]: [K, N] replica; C[g]: [M, N] full result
        buffer on GPU g.  Asynchronous: call synchronize() before reading C."""
        suf = _SUFFIX[C[0].dtype]
        M, N = C[0].shape
        K = B[0].shape[0]
        for g in range(self.ndev):
            r0, n = self.rows(M, g)
            if tuple(A_local[g].shape) != (n, K) or tuple(B[g].shape) != (K, N) or tuple(C[g].shape) != (M, N):
                raise IndexError(f"gemm_rowsharded: GPU {g} operands do not match the row partition ({n} rows)")
            if A_local[g].stride(1) != 1 or B[g].stride(1) != 1 or C[g].stride(1) != 1:
                raise ValueError("gemm_rowsharded: operands must be row-major (unit column stride)")
        P = ctypes.c_void_p * self.ndev
        pa = P(*[t.data_ptr() for t in A_local]); pb = P(*[t.data_ptr() for t in B]); pc = P(*[t.data_ptr() for t in C])
        ct = _capi.CTYPE[suf]
        _capi.check(getattr(_capi.lib(), f"am_mg_gemm_rowsharded_{suf}")(
            self._h, M, N, K, ct(alpha), pa, A_local[0].stride(0) if A_local[0].numel() else K, pb, B[0].stride(0), pc, C[0].stride(0)))

    def host_gemm_f32(self, alpha, A, B, C):
        """A [M,K], B [K,N], C [M,N]: row-major float32 HOST tensors (pinned for speed).  Synchronous."""
        M, K = A.shape
        N = B.shape[1]
        _capi.check(_capi.lib().am_mg_host_gemm_f32(self._h, M, N, K, float(alpha), A.data_ptr(), A.stride(0), B.data_ptr(),
                                                    B.stride(0), C.data_ptr(), C.stride(0)))
        return C

    def synchronize(self):
        _capi.check(_capi.lib().am_mg_synchronize(self._h))

    def close(self):
        if self._h:
            _capi.lib().am_mg_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
