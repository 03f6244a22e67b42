"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/am_b200.h declares, validates arguments before touching the device, and the host
package refuses to run without it (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from arraymancer_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "am_b200.h")).read()
    names = set(re.findall(r"\b(am_[a-z0-9_]+)\s*\(", src))
    # macro-declared families
    for suf in ("f32", "f64", "i32", "i64"):
        names.add(f"am_conv2d_forward_{suf}")
        names.add(f"am_conv2d_forward_act_{suf}")
        names.add(f"am_conv2d_backward_{suf}")
        names.add(f"am_conv2d_forward_strided_{suf}")            # AM_DECL_CONV_STRIDED(SUF, T)
        names.add(f"am_conv2d_backward_strided_{suf}")
        names.add(f"am_gemm_strided_batched_{suf}")              # AM_DECL_BATCHED(SUF, T)
        names.add(f"am_mg_gemm_rowsharded_{suf}")                # AM_DECL_MG(SUF, T)
    for suf in ("f32", "f64"):                                   # AM_DECL_NN(SUF, T)
        for op in _capi.NN_OPS:
            names.add(f"am_{op}_{suf}")
    names.discard("am_conv2d_forward_")
    names.discard("am_conv2d_forward_act_")
    names.discard("am_conv2d_backward_")
    return {n for n in names if not n.endswith("_")}


def test_library_exports_every_declared_symbol():
    lib = _capi.lib()
    declared = _header_symbols()
    assert declared, "no declarations parsed from include/am_b200.h"
    missing = sorted(n for n in declared if not hasattr(lib, n))
    assert not missing, f"declared in am_b200.h but not exported: {missing}"
    assert set(_capi.EXPORTED_SYMBOLS) <= declared


def test_version_and_error_string():
    assert "sm_100a" in _capi.version()
    lib = _capi.lib()
    assert lib.am_set_f32_path(99) == _capi.AM_ERR_INVALID
    assert b"selector" in lib.am_last_error()
    assert lib.am_set_f32_path(_capi.F32_AUTO) == _capi.AM_OK
    assert lib.am_get_f32_path() == _capi.F32_AUTO


def test_conv_out_dims_host_side():
    lib = _capi.lib()
    ho, wo = ctypes.c_int64(), ctypes.c_int64()
    # LeNet cv1: 28x28, 5x5, pad 0, stride 1 -> 24x24 (ex02_mnist.nim:25)
    d = _capi.ConvDesc(4096, 1, 28, 28, 20, 5, 5, 0, 0, 1, 1, 1, 1)
    assert lib.am_conv2d_out_dims(ctypes.byref(d), ctypes.byref(ho), ctypes.byref(wo)) == 0
    assert (ho.value, wo.value) == (24, 24)
    # strided 5x5 / 3x3 / pad 1 / stride 2 -> 3x3 (test_nnp_convolution.nim:54-134): the CPU formula,
    # not cuDNN helper's stride>1 precedence bug
    d = _capi.ConvDesc(1, 3, 5, 5, 2, 3, 3, 1, 1, 2, 2, 1, 1)
    lib.am_conv2d_out_dims(ctypes.byref(d), ctypes.byref(ho), ctypes.byref(wo))
    assert (ho.value, wo.value) == (3, 3)
    assert ctypes.sizeof(_capi.ConvDesc) == 13 * 8


def test_argument_validation_happens_before_any_device_work():
    lib = _capi.lib()
    # negative dimension
    rc = lib.am_gemm_strided_f64(None, -1, 2, 2, 1.0, None, 1, 1, None, 1, 1, 0.0, None, 1, 1)
    assert rc == _capi.AM_ERR_INVALID
    # null pointers with a non-empty shape
    rc = lib.am_gemm_strided_i64(None, 2, 2, 2, 1, None, 2, 1, None, 2, 1, 0, None, 2, 1)
    assert rc == _capi.AM_ERR_INVALID and b"null" in lib.am_last_error()
    # empty problems are no-ops (K == 0 leaves C untouched, gemm.nim:203)
    assert lib.am_gemm_strided_f32(None, 0, 5, 5, 1.0, None, 1, 1, None, 1, 1, 0.0, None, 1, 1) == 0
    assert lib.am_gemm_strided_f32(None, 5, 5, 0, 1.0, None, 1, 1, None, 1, 1, 2.0, None, 1, 1) == 0
    # cuBLAS-shaped adapter: bad op / leading dimension
    assert lib.am_cublas_gemm_f32(None, 2, 0, 4, 4, 4, 1.0, None, 4, None, 4, 0.0, None, 4) == _capi.AM_ERR_INVALID
    assert lib.am_cublas_gemm_f64(None, 0, 0, 4, 4, 4, 1.0, None, 3, None, 4, 0.0, None, 4) == _capi.AM_ERR_NONCONTIGUOUS
    # conv: null descriptor / bad geometry
    assert lib.am_conv2d_forward_f32(None, None, None, None, None, None) == _capi.AM_ERR_INVALID
    d = _capi.ConvDesc(1, 1, 2, 2, 1, 5, 5, 0, 0, 1, 1, 1, 1)   # kernel larger than the image
    assert lib.am_conv2d_forward_f32(None, ctypes.byref(d), None, None, None, None) == _capi.AM_ERR_INVALID


def test_no_cpu_fallback_when_library_is_missing(monkeypatch):
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", os.path.join(ROOT, "arraymancer_b200", "does_not_exist.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _capi.lib()


def test_host_ops_refuse_cpu_tensors():
    import torch
    import arraymancer_b200 as am
    a = torch.zeros(2, 2)
    with pytest.raises(ValueError, match="GPU"):
        am.gemm_strided(1, a, a, 0, a)
    with pytest.raises(ValueError, match="GPU"):
        am.conv2d(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 3, 3), None)


def test_host_shape_errors_match_reference_conventions():
    import torch
    import arraymancer_b200 as am
    a = torch.zeros(2, 3)
    with pytest.raises(IndexError):          # check_matmat -> IndexDefect (p_checks.nim:159-167)
        am.gemm_strided(1, a, a, 0, torch.zeros(2, 3))
    with pytest.raises(ValueError):          # rank mismatch -> ValueError (operators_blas_l2l3_cuda.nim:87)
        am.gemm_strided(1, torch.zeros(2), a, 0, a)
    assert am.conv_out_dims((4096, 20, 12, 12), (50, 20, 5, 5)) == (4096, 50, 8, 8)
