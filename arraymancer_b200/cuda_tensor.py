"""Host-side mirror of the reference's CudaTensor operator interface for the dense-contraction
path.  Same names, argument meaning and error behaviour as

  * tensor/data_structure.nim:44-58          CudaTensor[T] (shape, strides, offset, storage)
  * tensor/init_cuda.nim:23-59               `.cuda()` (column-major, H2D) / `.cpu()` (D2H)
  * tensor/operators_blas_l2l3_cuda.nim:43-87  `*`, cudaMM_C_eq_aAB_p_bC
  * tensor/operators_blas_l2l3.nim:58-100    `gemm(alpha, A, B, beta, C)`
  * laser/.../gemm.nim:192-201               `gemm_strided` (raw views)

(paths relative to /root/reference/src/arraymancer/).  Nim is not available in this image, so
this thin Python layer stands where the Nim `{.importc.}` bindings of INTEGRATION.md would;
torch is used only for device memory and streams.  All arithmetic happens in
libarraymancer_b200.so — there is no CPU fallback.

Differences from the reference, all supersets: operands may be arbitrary strided views
(the reference raises ValueError for non-contiguous CUDA operands, operators_blas_l2l3_cuda.nim:49-50,
marked TODO there) and the element type may be int32/int64 as well as float32/float64
(reference: SomeFloat only, data_structure.nim:44).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _capi

_SUFFIX = {torch.float32: "f32", torch.float64: "f64", torch.int32: "i32", torch.int64: "i64"}
_NP2T = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
         np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def gemm_strided(alpha, A: torch.Tensor, B: torch.Tensor, beta, C: torch.Tensor) -> torch.Tensor:
    """C <- alpha*A@B + beta*C on 2-D CUDA views of any stride (laser gemm_strided, gemm.nim:192-201).

    A, B, C are torch CUDA tensors used purely as (device pointer, strides) carriers — exactly
    what the Nim side passes as (get_offset_ptr, strides[0], strides[1])."""
    if A.dim() != 2 or B.dim() != 2 or C.dim() != 2:
        raise ValueError("gemm_strided: operands must be rank-2")
    if A.dtype not in _SUFFIX or B.dtype != A.dtype or C.dtype != A.dtype:
        raise TypeError("gemm_strided: operands must share one of float32/float64/int32/int64")
    M, K = A.shape
    K2, N = B.shape
    if K != K2 or tuple(C.shape) != (M, N):
        # check_matmat (tensor/private/p_checks.nim:159-167) raises IndexDefect
        raise IndexError(f"gemm_strided: shape mismatch {tuple(A.shape)} * {tuple(B.shape)} -> {tuple(C.shape)}")
    if not (A.is_cuda and B.is_cuda and C.is_cuda):
        raise ValueError("gemm_strided: operands must live on the GPU (no CPU fallback)")
    _gemm_raw(_SUFFIX[A.dtype], C.device, M, N, K, alpha, (A.data_ptr(), A.stride(0), A.stride(1)),
              (B.data_ptr(), B.stride(0), B.stride(1)), beta, (C.data_ptr(), C.stride(0), C.stride(1)))
    return C


def _gemm_raw(suf, device, M, N, K, alpha, a, b, beta, c) -> None:
    """The C-ABI call itself: a, b, c are (device pointer to the first logical element, rowStride,
    colStride) — what Nim passes as (get_offset_ptr, strides[0], strides[1]); strides may be negative."""
    ct = _capi.CTYPE[suf]
    with torch.cuda.device(device):
        _capi.check(getattr(_capi.lib(), f"am_gemm_strided_{suf}")(
            torch.cuda.current_stream(device).cuda_stream, M, N, K, ct(alpha), a[0], a[1], a[2],
            b[0], b[1], b[2], ct(beta), c[0], c[1], c[2]))


def gemm_strided_batched(alpha, A: torch.Tensor, B: torch.Tensor, beta, C: torch.Tensor) -> torch.Tensor:
    """C[b] <- alpha*A[b]@B[b] + beta*C[b] for rank-3 CUDA views [batch, rows, cols] of any strides (a batch stride of 0,
    e.g. an `expand`ed operand, shares that operand between the products): am_gemm_strided_batched_* — the device
    counterpart of `cublas_gemmStridedBatched` (tensor/backend/cublas.nim:172-208)."""
    if A.dim() != 3 or B.dim() != 3 or C.dim() != 3:
        raise ValueError("gemm_strided_batched: operands must be rank-3 [batch, rows, cols]")
    if A.dtype not in _SUFFIX or B.dtype != A.dtype or C.dtype != A.dtype:
        raise TypeError("gemm_strided_batched: operands must share one of float32/float64/int32/int64")
    nb, M, K = A.shape
    nb2, K2, N = B.shape
    if nb != nb2 or K != K2 or tuple(C.shape) != (nb, M, N):
        raise IndexError(f"gemm_strided_batched: shape mismatch {tuple(A.shape)} * {tuple(B.shape)} -> {tuple(C.shape)}")
    if not (A.is_cuda and B.is_cuda and C.is_cuda):
        raise ValueError("gemm_strided_batched: operands must live on the GPU (no CPU fallback)")
    suf = _SUFFIX[A.dtype]
    ct = _capi.CTYPE[suf]
    with torch.cuda.device(C.device):
        _capi.check(getattr(_capi.lib(), f"am_gemm_strided_batched_{suf}")(
            _stream_ptr(C), nb, M, N, K, ct(alpha), A.data_ptr(), A.stride(1), A.stride(2), A.stride(0),
            B.data_ptr(), B.stride(1), B.stride(2), B.stride(0), ct(beta), C.data_ptr(), C.stride(1), C.stride(2), C.stride(0)))
    return C


def cublas_gemm(transa: int, transb: int, m: int, n: int, k: int, alpha, A: torch.Tensor, lda: int,
                B: torch.Tensor, ldb: int, beta, C: torch.Tensor, ldc: int) -> None:
    """Column-major cuBLAS-shaped entry (tensor/backend/cublas.nim:142-170); A, B, C are flat
    device buffers."""
    suf = _SUFFIX[A.dtype]
    if suf not in ("f32", "f64"):
        raise TypeError("cublas_gemm: float32/float64 only (cublas.nim:142 `T: SomeFloat`)")
    ct = _capi.CTYPE[suf]
    with torch.cuda.device(C.device):
        _capi.check(getattr(_capi.lib(), f"am_cublas_gemm_{suf}")(
            _stream_ptr(C), transa, transb, m, n, k, ct(alpha), A.data_ptr(), lda, B.data_ptr(), ldb,
            ct(beta), C.data_ptr(), ldc))


class CudaTensor:
    """Mirror of CudaTensor[T] (tensor/data_structure.nim:44-58): shape, strides (in elements),
    offset, reference-counted device storage.  Default layout is COLUMN-major like the
    reference (tensor/private/p_init_cuda.nim:29-46)."""

    __slots__ = ("storage", "shape", "strides", "offset")

    def __init__(self, storage: torch.Tensor, shape, strides, offset: int = 0):
        self.storage = storage            # flat 1-D torch CUDA tensor (owns the cudaMalloc'd block)
        self.shape = tuple(int(s) for s in shape)
        self.strides = tuple(int(s) for s in strides)
        self.offset = int(offset)

    # ---- construction / transfer (init_cuda.nim:23-59)
    @staticmethod
    def new(shape, dtype=torch.float32, device="cuda", layout="colMajor") -> "CudaTensor":
        """newCudaTensor: uninitialised device tensor (p_init_cuda.nim:19-46)."""
        shape = tuple(int(s) for s in shape)
        n = int(np.prod(shape)) if shape else 1
        storage = torch.empty(n, dtype=dtype, device=device)
        return CudaTensor(storage, shape, _strides_for(shape, layout), 0)

    @property
    def dtype(self):
        return self.storage.dtype

    @property
    def rank(self) -> int:
        return len(self.shape)

    def view(self) -> torch.Tensor:
        """The strided torch view (only for non-negative strides; torch cannot express the others)."""
        return torch.as_strided(self.storage, self.shape, self.strides, self.offset)

    def offset_ptr(self) -> int:
        """get_offset_ptr (tensor/data_structure.nim:193-198): device address of the first logical element."""
        return self.storage.data_ptr() + self.offset * self.storage.element_size()

    def cpu(self) -> np.ndarray:
        """Blocking D2H copy of the whole storage, keeping the strides (init_cuda.nim:43-59)."""
        host = self.storage.cpu().numpy()
        it = host.dtype.itemsize
        return np.lib.stride_tricks.as_strided(host[self.offset:], self.shape, tuple(s * it for s in self.strides))

    # ---- views (tensor/shapeshifting_cuda.nim: transpose is a stride swap, no copy)
    def transpose(self) -> "CudaTensor":
        return CudaTensor(self.storage, self.shape[::-1], self.strides[::-1], self.offset)

    def __getitem__(self, idx) -> "CudaTensor":
        """Basic slicing with steps (negative allowed): offset += a*stride, stride *= step
        (tensor/private/p_accessors_macros_read.nim:54-58)."""
        if not isinstance(idx, tuple):
            idx = (idx,)
        shape, strides, off = [], [], self.offset
        for d, s in enumerate(idx):
            if not isinstance(s, slice):
                raise TypeError("CudaTensor slicing takes slices only")
            start, stop, step = s.indices(self.shape[d])
            n = len(range(start, stop, step))
            off += start * self.strides[d]
            shape.append(n)
            strides.append(self.strides[d] * step)
        for d in range(len(idx), self.rank):
            shape.append(self.shape[d])
            strides.append(self.strides[d])
        return CudaTensor(self.storage, shape, strides, off)

    # ---- `*` (operators_blas_l2l3_cuda.nim:74-87)
    def __mul__(self, other: "CudaTensor") -> "CudaTensor":
        return matmul(self, other)

    __matmul__ = __mul__


def _strides_for(shape, layout):
    if layout == "colMajor":
        st, acc = [], 1
        for s in shape:
            st.append(acc)
            acc *= s
        return tuple(st)
    st, acc = [], 1
    for s in reversed(shape):
        st.append(acc)
        acc *= s
    return tuple(reversed(st))


def cuda(t, device="cuda") -> CudaTensor:
    """`t.cuda()`: asContiguous(colMajor, force=true) then H2D (init_cuda.nim:23-41)."""
    arr = np.asarray(t)
    if arr.dtype not in _NP2T:
        raise TypeError(f"unsupported element type {arr.dtype}")
    flat = np.ascontiguousarray(arr.T).reshape(-1)       # column-major element order
    storage = torch.from_numpy(flat).to(device, non_blocking=False)
    return CudaTensor(storage, arr.shape, _strides_for(arr.shape, "colMajor"), 0)


def gemm(alpha, A: CudaTensor, B: CudaTensor, beta, C: CudaTensor) -> None:
    """C = alpha*A*B + beta*C (tensor/operators_blas_l2l3.nim:58-81; cudaMM_C_eq_aAB_p_bC)."""
    if A.rank != 2 or B.rank != 2 or C.rank != 2:
        raise ValueError("gemm: inputs must be matrices")
    if A.dtype not in _SUFFIX or B.dtype != A.dtype or C.dtype != A.dtype:
        raise TypeError("gemm: operands must share one of float32/float64/int32/int64")
    M, K = A.shape
    K2, N = B.shape
    if K != K2 or C.shape != (M, N):
        raise IndexError(f"gemm: shape mismatch {A.shape} * {B.shape} -> {C.shape}")   # check_matmat
    _gemm_raw(_SUFFIX[A.dtype], C.storage.device, M, N, K, alpha, (A.offset_ptr(), A.strides[0], A.strides[1]),
              (B.offset_ptr(), B.strides[0], B.strides[1]), beta, (C.offset_ptr(), C.strides[0], C.strides[1]))


def matmul(a: CudaTensor, b: CudaTensor) -> CudaTensor:
    """`a * b` for CudaTensor (operators_blas_l2l3_cuda.nim:74-87): rank-2 x rank-2; result is
    a fresh column-major tensor, alpha = 1, beta = 0 (never reads the uninitialised C)."""
    if a.rank == 2 and b.rank == 2:
        if a.shape[1] != b.shape[0]:
            raise IndexError(f"matmul: inner dimensions differ: {a.shape} * {b.shape}")   # check_matmat
        out = CudaTensor.new((a.shape[0], b.shape[1]), a.dtype, a.storage.device)
        if a.shape[1] == 0:
            out.storage.zero_()      # K == 0: the kernel leaves C untouched (gemm.nim:203); an empty sum is 0, like cuBLAS with beta = 0
        gemm(1, a, b, 0, out)
        return out
    if a.rank == 2 and b.rank == 1:
        if a.shape[1] != b.shape[0]:
            raise IndexError(f"matmul: inner dimensions differ: {a.shape} * {b.shape}")   # check_matvec
        # matrix-vector: a GEMM with N = 1 (the reference calls cublas gemv; GEMV is outside the
        # accelerated path, SURVEY §2.1 — served by the same strided kernel)
        bv = CudaTensor(b.storage, (b.shape[0], 1), (b.strides[0], 1), b.offset)
        out = CudaTensor.new((a.shape[0],), a.dtype, a.storage.device)
        ov = CudaTensor(out.storage, (a.shape[0], 1), (1, a.shape[0]), 0)
        if a.shape[1] == 0:
            out.storage.zero_()
        gemm(1, a, bv, 0, ov)
        return out
    raise ValueError("Matrix-Matrix or Matrix-Vector multiplication valid only if first Tensor is a Matrix "
                     "and second is a Matrix or Vector")


class PackedF32:
    """Pre-packed float32 operand for the tcgen05 3xTF32 path (am_pack_f32_a / am_pack_f32_b;
    device counterpart of laser's gemm_prepackA/B, gemm_prepacked.nim:178-270): split once into
    two K-major tf32 planes, reuse over many products."""

    def __init__(self, t: torch.Tensor, role: str):
        if t.dim() != 2 or t.dtype != torch.float32 or not t.is_cuda:
            raise ValueError("PackedF32: rank-2 float32 CUDA tensor expected")
        if role not in ("a", "b"):
            raise ValueError("PackedF32: role must be 'a' (A[M,K]) or 'b' (B[K,N])")
        self.role, self.shape, self.device = role, tuple(t.shape), t.device
        self._h = ctypes.c_void_p()
        fn = getattr(_capi.lib(), f"am_pack_f32_{role}")
        with torch.cuda.device(t.device):
            _capi.check(fn(_stream_ptr(t), t.shape[0], t.shape[1], t.data_ptr(), t.stride(0), t.stride(1),
                           ctypes.byref(self._h)))

    def repack(self, t: torch.Tensor) -> None:
        if tuple(t.shape) != self.shape or t.dtype != torch.float32:
            raise ValueError("PackedF32.repack: shape / dtype must match the original operand")
        fn = getattr(_capi.lib(), f"am_repack_f32_{self.role}")
        with torch.cuda.device(t.device):
            _capi.check(fn(_stream_ptr(t), self._h, t.data_ptr(), t.stride(0), t.stride(1)))

    def free(self) -> None:
        if self._h:
            _capi.lib().am_packed_free_f32(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def gemm_packed_bcast(alpha: float, A: PackedF32, B: PackedF32, C_view: torch.Tensor, peer_ptrs, self_index: int) -> torch.Tensor:
    """C_view <- alpha*A*B, stored by the GEMM epilogue to the same block of EVERY GPU's copy of C
    (row-sharded GEMM fused with the all-gather of its result).  `peer_ptrs[g]` is the
    address, mapped into this process (symmetric / peer memory), of C_view's first element in GPU
    g's buffer; `peer_ptrs[self_index]` is this GPU's own address (== C_view.data_ptr()).  The caller orders the GPUs (stream order +
    a cross-GPU barrier) before any of them reads C."""
    if A.role != "a" or B.role != "b" or A.shape[1] != B.shape[0] or tuple(C_view.shape) != (A.shape[0], B.shape[1]):
        raise IndexError("gemm_packed_bcast: operand roles / shapes do not match")
    n = len(peer_ptrs)
    if not 1 <= n <= 8:
        raise ValueError("gemm_packed_bcast: 1..8 peers")
    arr = (ctypes.c_void_p * n)(*[ctypes.c_void_p(int(x)) for x in peer_ptrs])
    with torch.cuda.device(C_view.device):
        _capi.check(_capi.lib().am_gemm_packed_f32_bcast(_stream_ptr(C_view), alpha, A._h, B._h, n, arr, int(self_index),
                                                         C_view.stride(0), C_view.stride(1)))
    return C_view


def gemm_packed(alpha: float, A: PackedF32, B: PackedF32, beta: float, C: torch.Tensor) -> torch.Tensor:
    """C <- alpha*A*B + beta*C from pre-packed operands (gemm_prepacked.nim:276-293 `gemm_packed`)."""
    if A.role != "a" or B.role != "b" or A.shape[1] != B.shape[0] or tuple(C.shape) != (A.shape[0], B.shape[1]):
        raise IndexError("gemm_packed: operand roles / shapes do not match")
    with torch.cuda.device(C.device):
        _capi.check(_capi.lib().am_gemm_packed_f32(_stream_ptr(C), alpha, A._h, B._h, beta, C.data_ptr(),
                                                   C.stride(0), C.stride(1)))
    return C
