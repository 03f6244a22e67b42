"""ORACLE — TEST INFRASTRUCTURE ONLY (numpy/ctypes front-end of oracle/liblaser_oracle.so).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker / the
reported CPU baseline.  The product package (``arraymancer_b200``) never imports it.

The shared library restates the reference's CPU algorithm
(laser ``gemm_strided``: /root/reference/src/arraymancer/laser/primitives/
matrix_multiplication/gemm.nim:192-273, and the im2col+GEMM conv:
nn_primitives/fallback/conv.nim:18-140).  Parity is pinned by the reference's own
known-answer vectors, see tests/golden/known_answers.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblaser_oracle.so")

# ISA variants of the reference's dispatch (laser_gemm.hpp::Variant)
DEFAULT_BUILD, AVX512, GENERIC, SSE_NO_FMA = 0, 1, 2, 3

_SUFFIX = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64",
           np.dtype(np.int32): "i32", np.dtype(np.int64): "i64"}
_CT = {"f32": ctypes.c_float, "f64": ctypes.c_double, "i32": ctypes.c_int32, "i64": ctypes.c_int64}


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (g++ -O3 -fopenmp, x86-64-v3)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "laser_gemm.hpp", "conv_oracle.hpp", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       env={k: v for k, v in os.environ.items() if k not in ("CXX", "CC")})
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        i64, p, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
        for suf, ct in _CT.items():
            f = getattr(_lib, f"oracle_gemm_strided_{suf}")
            f.argtypes = [i64, i64, i64, ct, p, i64, i64, p, i64, i64, ct, p, i64, i64, ci, ci]
            f.restype = None
            getattr(_lib, f"oracle_im2col_{suf}").argtypes = [p, p, p]
            getattr(_lib, f"oracle_col2im_{suf}").argtypes = [p, p, p]
            getattr(_lib, f"oracle_conv2d_forward_{suf}").argtypes = [p, p, p, p, p, ci, ci]
            getattr(_lib, f"oracle_conv2d_backward_{suf}").argtypes = [p, p, p, p, p, p, p, ci, ci]
            for n in ("im2col", "col2im", "conv2d_forward", "conv2d_backward"):
                getattr(_lib, f"oracle_{n}_{suf}").restype = None
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


class _NativeGemm:
    """The same restatement compiled `-march=native` (bench.py's CPU arm only): lets the `-d:avx512` tile shapes
    (14x32 / 14x16, gemm_tiling.nim:89-109) run with real AVX-512 registers on hosts that have them."""

    def __init__(self, cdll):
        self._l = cdll
        for suf, ct in _CT.items():
            f = getattr(cdll, f"oracle_gemm_strided_{suf}")
            f.argtypes = [ctypes.c_int64] * 3 + [ct, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                                  ctypes.c_int64, ctypes.c_int64, ct, ctypes.c_void_p, ctypes.c_int64,
                                                  ctypes.c_int64, ctypes.c_int, ctypes.c_int]
            f.restype = None

    def gemm_strided(self, alpha, A, B, beta, C, variant: int = AVX512, threads: int = 0):
        suf = _SUFFIX[A.dtype]
        ct = _CT[suf]
        (rsA, csA), (rsB, csB), (rsC, csC) = _estrides(A), _estrides(B), _estrides(C)
        getattr(self._l, f"oracle_gemm_strided_{suf}")(A.shape[0], B.shape[1], A.shape[1], ct(alpha), A.ctypes.data, rsA, csA,
                                                       B.ctypes.data, rsB, csB, ct(beta), C.ctypes.data, rsC, csC, variant, threads)
        return C


_native = False


def native():
    """Build (once, on the machine that runs the CPU arm) and load oracle/liblaser_oracle_native.so; None when the host
    has no AVX-512 (the 14x32 tiles would only spill) or the build fails."""
    global _native
    if _native is not False:
        return _native
    _native = None
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        flags = ""
    if " avx512f" not in flags:
        return None
    import hashlib
    # -march=native code must never run on another CPU model: the file name carries a hash of this host's model + flags
    ident = "".join(l for l in flags.splitlines() if l.startswith(("model name", "flags")))[:8192]
    path = os.path.join(_HERE, f"liblaser_oracle_native_{hashlib.sha1(ident.encode()).hexdigest()[:10]}.so")
    src = os.path.join(_HERE, "oracle.cpp")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    deps = [os.path.join(_HERE, f) for f in ("oracle.cpp", "laser_gemm.hpp", "conv_oracle.hpp")]
    if not os.path.exists(path) or any(os.path.getmtime(d) > os.path.getmtime(path) for d in deps):
        subprocess.run([cxx, "-O3", "-std=c++17", "-fopenmp", "-march=native", "-ffp-contract=off", "-fPIC", "-fvisibility=hidden",
                        "-shared", "-o", path, src], check=True, env={k: v for k, v in os.environ.items() if k not in ("CXX", "CC")})
    _native = _NativeGemm(ctypes.CDLL(path))
    return _native


def _estrides(a: np.ndarray):
    assert a.ndim == 2
    it = a.dtype.itemsize
    assert a.strides[0] % it == 0 and a.strides[1] % it == 0
    return a.strides[0] // it, a.strides[1] // it


def gemm_strided(alpha, A: np.ndarray, B: np.ndarray, beta, C: np.ndarray,
                 variant: int = DEFAULT_BUILD, threads: int = 0) -> np.ndarray:
    """C <- alpha*A@B + beta*C in place, any 2-D numpy views (signed/zero strides allowed),
    exactly as laser gemm_strided would be called with (get_offset_ptr, strides[0], strides[1])."""
    suf = _SUFFIX[A.dtype]
    assert B.dtype == A.dtype and C.dtype == A.dtype
    M, K = A.shape
    K2, N = B.shape
    assert K == K2 and C.shape == (M, N)
    ct = _CT[suf]
    rsA, csA = _estrides(A)
    rsB, csB = _estrides(B)
    rsC, csC = _estrides(C)
    getattr(lib(), f"oracle_gemm_strided_{suf}")(
        M, N, K, ct(alpha), A.ctypes.data, rsA, csA, B.ctypes.data, rsB, csB,
        ct(beta), C.ctypes.data, rsC, csC, variant, threads)
    return C


def matmul(A: np.ndarray, B: np.ndarray, variant: int = DEFAULT_BUILD, threads: int = 0) -> np.ndarray:
    """`*` of tensor/operators_blas_l2l3.nim:88-100: row-major uninitialised C, alpha=1, beta=0."""
    C = np.empty((A.shape[0], B.shape[1]), dtype=A.dtype)
    return gemm_strided(1, A, B, 0, C, variant, threads)


def _dims(inp_shape, k_shape, padding, stride, dilation=(1, 1)):
    N, C, H, W = inp_shape
    Cout, C2, kH, kW = k_shape
    assert C == C2
    return np.array([N, C, H, W, Cout, kH, kW, padding[0], padding[1], stride[0], stride[1],
                     dilation[0], dilation[1]], dtype=np.int64)


def conv_out_hw(inp_shape, k_shape, padding, stride, dilation=(1, 1)):
    _, _, H, W = inp_shape
    _, _, kH, kW = k_shape
    Ho = (H + 2 * padding[0] - (dilation[0] * (kH - 1) + 1)) // stride[0] + 1
    Wo = (W + 2 * padding[1] - (dilation[1] * (kW - 1) + 1)) // stride[1] + 1
    return Ho, Wo


def im2col(image: np.ndarray, k_hw, padding=(0, 0), stride=(1, 1), dilation=(1, 1)) -> np.ndarray:
    image = np.ascontiguousarray(image)
    C, H, W = image.shape
    d = _dims((1, C, H, W), (1, C, k_hw[0], k_hw[1]), padding, stride, dilation)
    Ho, Wo = conv_out_hw((1, C, H, W), (1, C, k_hw[0], k_hw[1]), padding, stride, dilation)
    out = np.empty((C * k_hw[0] * k_hw[1], Ho * Wo), dtype=image.dtype)
    getattr(lib(), f"oracle_im2col_{_SUFFIX[image.dtype]}")(image.ctypes.data, d.ctypes.data, out.ctypes.data)
    return out


def col2im(cols: np.ndarray, chw, k_hw, padding=(0, 0), stride=(1, 1), dilation=(1, 1)) -> np.ndarray:
    cols = np.ascontiguousarray(cols)
    C, H, W = chw
    d = _dims((1, C, H, W), (1, C, k_hw[0], k_hw[1]), padding, stride, dilation)
    out = np.empty((C, H, W), dtype=cols.dtype)
    getattr(lib(), f"oracle_col2im_{_SUFFIX[cols.dtype]}")(cols.ctypes.data, d.ctypes.data, out.ctypes.data)
    return out


def conv2d(inp: np.ndarray, kernel: np.ndarray, bias: np.ndarray | None, padding=(0, 0), stride=(1, 1),
           dilation=(1, 1), variant: int = DEFAULT_BUILD, threads: int = 0) -> np.ndarray:
    """im2colgemm_conv2d (fallback/conv.nim:81-106)."""
    inp = np.ascontiguousarray(inp)
    kernel = np.ascontiguousarray(kernel)
    d = _dims(inp.shape, kernel.shape, padding, stride, dilation)
    Ho, Wo = conv_out_hw(inp.shape, kernel.shape, padding, stride, dilation)
    out = np.empty((inp.shape[0], kernel.shape[0], Ho, Wo), dtype=inp.dtype)
    b = None
    if bias is not None:
        b = np.ascontiguousarray(bias.reshape(-1)).astype(inp.dtype, copy=False)
        assert b.size == kernel.shape[0]
    getattr(lib(), f"oracle_conv2d_forward_{_SUFFIX[inp.dtype]}")(
        inp.ctypes.data, kernel.ctypes.data, b.ctypes.data if b is not None else None,
        out.ctypes.data, d.ctypes.data, variant, threads)
    return out


def conv2d_backward(inp: np.ndarray, kernel: np.ndarray, grad_output: np.ndarray, with_bias: bool = True,
                    padding=(0, 0), stride=(1, 1), dilation=(1, 1), variant: int = DEFAULT_BUILD,
                    threads: int = 0):
    """conv2d_backward (nnp_convolution.nim:65-107 -> fallback/conv.nim:108-140).
    Returns (grad_input, grad_weight, grad_bias[Cout,1,1] or None)."""
    inp = np.ascontiguousarray(inp)
    kernel = np.ascontiguousarray(kernel)
    grad_output = np.ascontiguousarray(grad_output)
    d = _dims(inp.shape, kernel.shape, padding, stride, dilation)
    gin = np.empty_like(inp)
    gw = np.empty_like(kernel)
    gb = np.empty((kernel.shape[0], 1, 1), dtype=inp.dtype) if with_bias else None
    getattr(lib(), f"oracle_conv2d_backward_{_SUFFIX[inp.dtype]}")(
        inp.ctypes.data, kernel.ctypes.data, grad_output.ctypes.data, gin.ctypes.data, gw.ctypes.data,
        gb.ctypes.data if gb is not None else None, d.ctypes.data, variant, threads)
    return gin, gw, gb
