"""ctypes binding of libarraymancer_b200.so — the C-ABI declared in include/am_b200.h.

The product path has NO fallback: if the CUDA library is missing or a call fails, this module
raises.  (The CPU oracle under oracle/ is test infrastructure and is never imported here.)
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AM_B200_LIB: developer override, e.g. the profiling build libarraymancer_b200_prof.so (make -C arraymancer_b200/csrc prof)
LIB_PATH = os.environ.get("AM_B200_LIB") or os.path.join(_HERE, "libarraymancer_b200.so")

AM_OK, AM_ERR_INVALID, AM_ERR_CUDA, AM_ERR_UNSUPPORTED, AM_ERR_NONCONTIGUOUS = range(5)
F32_AUTO, F32_SIMT, F32_TC, F32_TC_1CTA = range(4)
F64_AUTO, F64_SIMT, F64_DMMA = range(3)

SUFFIXES = ("f32", "f64", "i32", "i64")
CTYPE = {"f32": ctypes.c_float, "f64": ctypes.c_double, "i32": ctypes.c_int32, "i64": ctypes.c_int64}

# every symbol include/am_b200.h declares (tests check the library exports all of them)
NN_OPS = ("relu_forward", "relu_backward", "maxpool2d_forward", "maxpool2d_backward", "maxpool2d_backward_relu", "linear_forward", "linear_backward",
          "sparse_softmax_cross_entropy", "sparse_softmax_cross_entropy_backward")
EXPORTED_SYMBOLS = (
    ["am_version", "am_last_error", "am_device_info", "am_shutdown", "am_set_f32_path", "am_get_f32_path", "am_set_f64_path", "am_get_f64_path", "am_set_conv_path",
     "am_cublas_gemm_f32", "am_cublas_gemm_f64", "am_pack_f32_a", "am_pack_f32_b", "am_repack_f32_a",
     "am_repack_f32_b", "am_gemm_packed_f32", "am_gemm_packed_f32_bcast", "am_packed_free_f32", "am_conv2d_out_dims", "am_kernel_launch_count", "am_microbench",
     "am_set_tuning", "am_get_tuning", "am_memcpy2d_async", "am_mg_init", "am_mg_destroy", "am_mg_device_count",
     "am_mg_stream", "am_mg_rows", "am_mg_synchronize", "am_mg_host_gemm_f32", "am_packed_floats_f32", "am_pack_f32_a_into", "am_pack_f32_b_into", "am_packed_wrap_f32"]
    + [f"am_gemm_strided_{s}" for s in SUFFIXES]
    + [f"am_host_gemm_strided_{s}" for s in SUFFIXES]
    + [f"am_conv2d_forward_{s}" for s in SUFFIXES]
    + [f"am_conv2d_forward_act_{s}" for s in SUFFIXES]
    + [f"am_conv2d_backward_{s}" for s in SUFFIXES]
    + [f"am_conv2d_forward_strided_{s}" for s in SUFFIXES]
    + [f"am_conv2d_backward_strided_{s}" for s in SUFFIXES]
    + [f"am_gemm_strided_batched_{s}" for s in SUFFIXES]
    + [f"am_mg_gemm_rowsharded_{s}" for s in SUFFIXES]
    + [f"am_{op}_{s}" for s in ("f32", "f64") for op in NN_OPS]
)


class ConvDesc(ctypes.Structure):
    """am_conv2d_desc"""
    _fields_ = [(n, ctypes.c_int64) for n in
                ("N", "C", "H", "W", "Cout", "kH", "kW", "padH", "padW", "strideH", "strideW", "dilH", "dilW")]


class AmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"arraymancer_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> ctypes.CDLL:
    """Load the CUDA library; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C arraymancer_b200/csrc`.  There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    i64, p, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
    L.am_version.restype = ctypes.c_char_p
    L.am_last_error.restype = ctypes.c_char_p
    L.am_device_info.argtypes = [ctypes.POINTER(ci)] * 3
    L.am_set_f32_path.argtypes = [ci]
    L.am_set_f64_path.argtypes = [ci]
    L.am_set_conv_path.argtypes = [ci]
    L.am_kernel_launch_count.restype = i64
    L.am_microbench.argtypes = [ci, ctypes.POINTER(ctypes.c_double)]
    L.am_conv2d_out_dims.argtypes = [ctypes.POINTER(ConvDesc), ctypes.POINTER(i64), ctypes.POINTER(i64)]
    for s in SUFFIXES:
        ct = CTYPE[s]
        getattr(L, f"am_gemm_strided_{s}").argtypes = [p, i64, i64, i64, ct, p, i64, i64, p, i64, i64, ct, p, i64, i64]
        getattr(L, f"am_host_gemm_strided_{s}").argtypes = [i64, i64, i64, ct, p, i64, i64, p, i64, i64, ct, p, i64, i64]
        getattr(L, f"am_conv2d_forward_{s}").argtypes = [p, ctypes.POINTER(ConvDesc), p, p, p, p]
        getattr(L, f"am_conv2d_forward_act_{s}").argtypes = [p, ctypes.POINTER(ConvDesc), p, p, p, p, ci]
        getattr(L, f"am_conv2d_backward_{s}").argtypes = [p, ctypes.POINTER(ConvDesc), p, p, p, p, p, p]
        getattr(L, f"am_conv2d_forward_strided_{s}").argtypes = [p, ctypes.POINTER(ConvDesc), p, p, p, p, p, i64, p, p, ci]
        getattr(L, f"am_conv2d_backward_strided_{s}").argtypes = [p, ctypes.POINTER(ConvDesc)] + [p] * 11 + [i64]
        getattr(L, f"am_gemm_strided_batched_{s}").argtypes = [p, i64, i64, i64, i64, ct, p, i64, i64, i64, p, i64, i64, i64,
                                                               ct, p, i64, i64, i64]
    f = ctypes.c_float
    L.am_pack_f32_a.argtypes = [p, i64, i64, p, i64, i64, ctypes.POINTER(p)]
    L.am_pack_f32_b.argtypes = [p, i64, i64, p, i64, i64, ctypes.POINTER(p)]
    L.am_repack_f32_a.argtypes = [p, p, p, i64, i64]
    L.am_repack_f32_b.argtypes = [p, p, p, i64, i64]
    L.am_gemm_packed_f32.argtypes = [p, f, p, p, f, p, i64, i64]
    L.am_gemm_packed_f32_bcast.argtypes = [p, f, p, p, ci, p, ci, i64, i64]
    L.am_packed_free_f32.argtypes = [p]
    L.am_packed_floats_f32.argtypes = [i64, i64]
    L.am_packed_floats_f32.restype = i64
    L.am_pack_f32_a_into.argtypes = [p, i64, i64, p, i64, i64, p, ctypes.POINTER(p)]
    L.am_pack_f32_b_into.argtypes = [p, i64, i64, p, i64, i64, p, ctypes.POINTER(p)]
    L.am_packed_wrap_f32.argtypes = [i64, i64, p, ctypes.POINTER(p)]
    L.am_mg_init.argtypes = [ci, ctypes.POINTER(ci), ctypes.POINTER(p)]
    L.am_mg_destroy.argtypes = [p]
    L.am_mg_device_count.argtypes = [p]
    L.am_mg_stream.argtypes = [p, ci]
    L.am_mg_stream.restype = p
    L.am_mg_rows.argtypes = [p, i64, ci, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    L.am_mg_synchronize.argtypes = [p]
    L.am_mg_host_gemm_f32.argtypes = [p, i64, i64, i64, f, p, i64, p, i64, p, i64]
    for s in SUFFIXES:
        getattr(L, f"am_mg_gemm_rowsharded_{s}").argtypes = [p, i64, i64, i64, CTYPE[s], p, i64, p, i64, p, i64]
    L.am_memcpy2d_async.argtypes = [p, p, i64, p, i64, i64, i64, ci]
    L.am_set_tuning.argtypes = [ctypes.c_char_p, ci]
    L.am_get_tuning.argtypes = [ctypes.c_char_p, ctypes.POINTER(ci)]
    for s in ("f32", "f64"):
        ct = CTYPE[s]
        getattr(L, f"am_cublas_gemm_{s}").argtypes = [p, ci, ci, i64, i64, i64, ct, p, i64, p, i64, ct, p, i64]
        getattr(L, f"am_relu_forward_{s}").argtypes = [p, i64, p, p]
        getattr(L, f"am_relu_backward_{s}").argtypes = [p, i64, p, p, p]
        getattr(L, f"am_maxpool2d_forward_{s}").argtypes = [p] + [i64] * 10 + [p, p, p]
        getattr(L, f"am_maxpool2d_backward_{s}").argtypes = [p, i64, i64, p, p, p, ci]
        getattr(L, f"am_maxpool2d_backward_relu_{s}").argtypes = [p, i64, i64, p, p, p, p, ci]
        getattr(L, f"am_linear_forward_{s}").argtypes = [p, i64, i64, i64, p, p, p, p]
        getattr(L, f"am_linear_backward_{s}").argtypes = [p, i64, i64, i64, p, p, p, p, p, p]
        getattr(L, f"am_sparse_softmax_cross_entropy_{s}").argtypes = [p, i64, i64, p, i64, i64, p, p]
        getattr(L, f"am_sparse_softmax_cross_entropy_backward_{s}").argtypes = [p, i64, i64, ct, p, i64, i64, p, p]
    _lib = L
    return L


def check(code: int) -> None:
    """Mirror of nimcuda's `check`: turn a non-zero status into an exception."""
    if code != AM_OK:
        raise AmError(code, lib().am_last_error().decode(errors="replace"))


def version() -> str:
    return lib().am_version().decode()


def device_info():
    a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(lib().am_device_info(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
    return a.value, b.value, c.value


def set_f32_path(path: int) -> None:
    check(lib().am_set_f32_path(path))


def set_f64_path(path: int) -> None:
    check(lib().am_set_f64_path(path))


CONV_AUTO, CONV_GATHER, CONV_DIRECT, CONV_TC = 0, 1, 2, 3
ACT_NONE, ACT_RELU = 0, 1


def set_conv_path(path: int) -> None:
    check(lib().am_set_conv_path(path))


def set_tuning(name: str, value: int) -> None:
    """am_set_tuning: explicit process-wide tuning knob (the library reads no environment variables)."""
    check(lib().am_set_tuning(name.encode(), int(value)))


def get_tuning(name: str) -> int:
    v = ctypes.c_int()
    check(lib().am_get_tuning(name.encode(), ctypes.byref(v)))
    return v.value


def kernel_launch_count() -> int:
    return int(lib().am_kernel_launch_count())


def microbench(which: int) -> float:
    out = ctypes.c_double()
    check(lib().am_microbench(which, ctypes.byref(out)))
    return out.value
