"""conv2d / conv2d_backward with the reference's signatures, on the fused implicit-GEMM kernels.

Mirrors (paths relative to /root/reference/src/arraymancer/):
  * nn_primitives/nnp_conv2d_cudnn.nim:20-72    conv2d(input, kernel, bias, padding, strides, dilation)
  * nn_primitives/nnp_conv2d_cudnn.nim:74-204   conv2d_backward(..., grad_output) -> grad_input, grad_kernel, grad_bias
  * nn_primitives/nnp_convolution.nim:28-107    the CPU dispatcher whose Im2ColGEMM branch
    (fallback/conv.nim:81-140) this replaces
Inputs are NCHW torch CUDA tensors (device memory carriers); the arithmetic is in
libarraymancer_b200.so.  The im2col buffer is never materialised.
"""
from __future__ import annotations

import ctypes

import torch

from . import _capi
from .cuda_tensor import _SUFFIX, _stream_ptr


def _desc(inp, kernel, padding, strides, dilation) -> _capi.ConvDesc:
    if inp.dim() != 4 or kernel.dim() != 4:
        raise ValueError("conv2d: input and kernel must be rank-4 (NCHW / [Cout,Cin,kH,kW])")
    N, C, H, W = inp.shape
    Cout, C2, kH, kW = kernel.shape
    if C != C2:
        raise IndexError(f"conv2d: input has {C} channels, kernel expects {C2}")
    return _capi.ConvDesc(N, C, H, W, Cout, kH, kW, padding[0], padding[1], strides[0], strides[1],
                          dilation[0], dilation[1])


def conv_out_dims(inp_shape, kernel_shape, padding=(0, 0), strides=(1, 1), dilation=(1, 1)):
    """convOutDims — with the arithmetic the CPU path uses (fallback/conv.nim:90-91); the cuDNN
    helper (backend/cudnn_conv_interface.nim:118-119) has a precedence bug for stride > 1
     that is deliberately not reproduced."""
    N, _, H, W = inp_shape
    Cout, _, kH, kW = kernel_shape
    Ho = (H + 2 * padding[0] - (dilation[0] * (kH - 1) + 1)) // strides[0] + 1
    Wo = (W + 2 * padding[1] - (dilation[1] * (kW - 1) + 1)) // strides[1] + 1
    return (N, Cout, Ho, Wo)


def _check_mem(*ts):
    dt = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise ValueError("conv2d: tensors must live on the GPU (no CPU fallback)")
        if not t.is_contiguous():
            raise ValueError("conv2d: tensors must be C-contiguous NCHW")
        dt = dt or t.dtype
        if t.dtype != dt or t.dtype not in _SUFFIX:
            raise TypeError("conv2d: tensors must share one of float32/float64/int32/int64")
    return _SUFFIX[dt]


def conv2d(input: torch.Tensor, kernel: torch.Tensor, bias: torch.Tensor | None = None,
           padding=(0, 0), strides=(1, 1), dilation=(1, 1)) -> torch.Tensor:
    """Cross-correlation of NCHW `input` with `kernel` [Cout,Cin,kH,kW] plus `bias` [Cout,1,1]
    (or None for the reference's rank-0 "no bias" tensor)."""
    d = _desc(input, kernel, padding, strides, dilation)
    if bias is not None and bias.numel() == 0:
        bias = None
    if bias is not None:
        if bias.numel() != kernel.shape[0]:
            raise IndexError("conv2d: bias must have Cout elements ([Cout,1,1])")
        bias = bias.reshape(-1)
    suf = _check_mem(input, kernel, bias)
    shape = conv_out_dims(input.shape, kernel.shape, padding, strides, dilation)
    if shape[2] <= 0 or shape[3] <= 0:
        raise ValueError("conv2d: kernel larger than the padded input")
    out = torch.empty(shape, dtype=input.dtype, device=input.device)
    with torch.cuda.device(input.device):
        _capi.check(getattr(_capi.lib(), f"am_conv2d_forward_{suf}")(
            _stream_ptr(input), ctypes.byref(d), input.data_ptr(), kernel.data_ptr(),
            bias.data_ptr() if bias is not None else None, out.data_ptr()))
    return out


def conv2d_backward(input: torch.Tensor, kernel: torch.Tensor, bias: torch.Tensor | None,
                    padding, strides, dilation, grad_output: torch.Tensor,
                    need_input_grad: bool = True, need_kernel_grad: bool = True):
    """Returns (grad_input, grad_kernel, grad_bias); grad_bias has bias' shape ([Cout,1,1]) or is
    None when there is no bias (nnp_convolution.nim:91-94).  `need_*_grad=False` skips that gradient
    (returned as None) — the C ABI takes NULL for any output it should not compute."""
    d = _desc(input, kernel, padding, strides, dilation)
    if bias is not None and bias.numel() == 0:
        bias = None
    suf = _check_mem(input, kernel, grad_output)
    want = conv_out_dims(input.shape, kernel.shape, padding, strides, dilation)
    if tuple(grad_output.shape) != tuple(want):
        raise IndexError(f"conv2d_backward: grad_output shape {tuple(grad_output.shape)} != {want}")
    gin = torch.empty_like(input) if need_input_grad else None
    gk = torch.empty_like(kernel) if need_kernel_grad else None
    gb = (torch.empty((kernel.shape[0], 1, 1), dtype=input.dtype, device=input.device)
          if (bias is not None and need_kernel_grad) else None)
    with torch.cuda.device(input.device):
        _capi.check(getattr(_capi.lib(), f"am_conv2d_backward_{suf}")(
            _stream_ptr(input), ctypes.byref(d), input.data_ptr(), kernel.data_ptr(), grad_output.data_ptr(),
            gin.data_ptr() if gin is not None else None, gk.data_ptr() if gk is not None else None,
            gb.data_ptr() if gb is not None else None))
    return gin, gk, gb
