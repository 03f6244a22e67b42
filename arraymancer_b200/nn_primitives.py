"""conv2d / conv2d_backward with the reference's signatures, on the fused implicit-GEMM kernels, and the
small operators around them (relu, maxpool2d, linear, sparse_softmax_cross_entropy: SURVEY 8f rows 1-3).

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
    """Returns (dtype suffix, all tensors dense NCHW?).  Non-dense views (transposed, sliced, Fortran-ordered ...) are
    legal: they cross the boundary with their 4 element strides (am_conv2d_*_strided_*), as the reference builds its
    descriptors from `t.strides` (nn_primitives/backend/cudnn.nim:59-75)."""
    dt, dense = None, True
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise ValueError("conv2d: tensors must live on the GPU (no CPU fallback)")
        dense = dense and t.is_contiguous()
        dt = dt or t.dtype
        if t.dtype != dt or t.dtype not in _SUFFIX:
            raise TypeError("conv2d: tensors must share one of float32/float64/int32/int64")
    return _SUFFIX[dt], dense


def _strides4(t):
    return (ctypes.c_int64 * 4)(*[int(x) for x in t.stride()])


def conv2d(input: torch.Tensor, kernel: torch.Tensor, bias: torch.Tensor | None = None,
           padding=(0, 0), strides=(1, 1), dilation=(1, 1), activation: str | None = None) -> torch.Tensor:
    """Cross-correlation of NCHW `input` with `kernel` [Cout,Cin,kH,kW] plus `bias` [Cout,1,1]
    (or None for the reference's rank-0 "no bias" tensor).  activation="relu" fuses the reference's separate
    `relu` pass (nnp_activation.nim:35-36) into the epilogue: max(0, conv + bias)."""
    if activation not in (None, "relu"):
        raise ValueError("conv2d: activation must be None or 'relu'")
    d = _desc(input, kernel, padding, strides, dilation)
    if bias is not None and bias.numel() == 0:
        bias = None
    if bias is not None:
        if bias.numel() != kernel.shape[0]:
            raise IndexError("conv2d: bias must have Cout elements ([Cout,1,1])")
        bias = bias.reshape(-1)
    suf, dense = _check_mem(input, kernel, bias)
    shape = conv_out_dims(input.shape, kernel.shape, padding, strides, dilation)
    if shape[2] <= 0 or shape[3] <= 0:
        raise ValueError("conv2d: kernel larger than the padded input")
    out = torch.empty(shape, dtype=input.dtype, device=input.device)
    with torch.cuda.device(input.device):
        if not dense:
            _capi.check(getattr(_capi.lib(), f"am_conv2d_forward_strided_{suf}")(
                _stream_ptr(input), ctypes.byref(d), input.data_ptr(), _strides4(input), kernel.data_ptr(), _strides4(kernel),
                bias.data_ptr() if bias is not None else None, int(bias.stride(0)) if bias is not None else 1,
                out.data_ptr(), None, _capi.ACT_RELU if activation else _capi.ACT_NONE))
        elif activation is None:
            _capi.check(getattr(_capi.lib(), f"am_conv2d_forward_{suf}")(
                _stream_ptr(input), ctypes.byref(d), input.data_ptr(), kernel.data_ptr(),
                bias.data_ptr() if bias is not None else None, out.data_ptr()))
        else:
            _capi.check(getattr(_capi.lib(), f"am_conv2d_forward_act_{suf}")(
                _stream_ptr(input), ctypes.byref(d), input.data_ptr(), kernel.data_ptr(),
                bias.data_ptr() if bias is not None else None, out.data_ptr(), _capi.ACT_RELU))
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
    suf, dense = _check_mem(input, kernel, grad_output)
    want = conv_out_dims(input.shape, kernel.shape, padding, strides, dilation)
    if tuple(grad_output.shape) != tuple(want):
        raise IndexError(f"conv2d_backward: grad_output shape {tuple(grad_output.shape)} != {want}")
    # fresh gradients are dense NCHW whatever the layout of the operands they belong to
    gin = torch.empty(input.shape, dtype=input.dtype, device=input.device) if need_input_grad else None
    gk = torch.empty(kernel.shape, dtype=kernel.dtype, device=kernel.device) if need_kernel_grad else None
    gb = (torch.empty((kernel.shape[0], 1, 1), dtype=input.dtype, device=input.device)
          if (bias is not None and need_kernel_grad) else None)
    if not dense:
        with torch.cuda.device(input.device):
            _capi.check(getattr(_capi.lib(), f"am_conv2d_backward_strided_{suf}")(
                _stream_ptr(input), ctypes.byref(d), input.data_ptr(), _strides4(input), kernel.data_ptr(), _strides4(kernel),
                grad_output.data_ptr(), _strides4(grad_output), gin.data_ptr() if gin is not None else None, None,
                gk.data_ptr() if gk is not None else None, None, gb.data_ptr() if gb is not None else None, 1))
        return gin, gk, gb
    with torch.cuda.device(input.device):
        _capi.check(getattr(_capi.lib(), f"am_conv2d_backward_{suf}")(
            _stream_ptr(input), ctypes.byref(d), input.data_ptr(), kernel.data_ptr(), grad_output.data_ptr(),
            gin.data_ptr() if gin is not None else None, gk.data_ptr() if gk is not None else None,
            gb.data_ptr() if gb is not None else None))
    return gin, gk, gb


# ---------------------------------------------------------------------------------------------
# LeNet companions.  Mirrors nn_primitives/nnp_activation.nim:35-36,65-70,
# nnp_maxpooling.nim:19-83, nnp_linear.nim:20-66, nnp_softmax_cross_entropy.nim:100-178,219-252.

_FSUF = {torch.float32: "f32", torch.float64: "f64"}


def _fcheck(name, *ts):
    dt = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise ValueError(f"{name}: tensors must live on the GPU (no CPU fallback)")
        if not t.is_contiguous():
            raise ValueError(f"{name}: tensors must be C-contiguous")
        dt = dt or t.dtype
        if t.dtype != dt or dt not in _FSUF:
            raise TypeError(f"{name}: tensors must share float32 or float64")
    return _FSUF[dt]


def relu(t: torch.Tensor) -> torch.Tensor:
    """max(0, x) element-wise (nnp_activation.nim:35-36)."""
    suf = _fcheck("relu", t)
    out = torch.empty_like(t)
    with torch.cuda.device(t.device):
        _capi.check(getattr(_capi.lib(), f"am_relu_forward_{suf}")(_stream_ptr(t), t.numel(), t.data_ptr(), out.data_ptr()))
    return out


def relu_backward(gradient: torch.Tensor, cached_tensor: torch.Tensor) -> torch.Tensor:
    """cached <= 0 ? 0 : gradient (nnp_activation.nim:65-70)."""
    suf = _fcheck("relu_backward", gradient, cached_tensor)
    if gradient.shape != cached_tensor.shape:
        raise IndexError("relu_backward: shapes differ")
    out = torch.empty_like(gradient)
    with torch.cuda.device(gradient.device):
        _capi.check(getattr(_capi.lib(), f"am_relu_backward_{suf}")(
            _stream_ptr(gradient), gradient.numel(), gradient.data_ptr(), cached_tensor.data_ptr(), out.data_ptr()))
    return out


def maxpool2d(input: torch.Tensor, kernel, padding=(0, 0), stride=(1, 1)):
    """-> (max_indices [N*C*outH*outW] int64 flat input indices, maxpooled [N,C,outH,outW])
    (nnp_maxpooling.nim:19-66)."""
    suf = _fcheck("maxpool2d", input)
    if input.dim() != 4:
        raise ValueError("maxpool2d: input must be rank-4 NCHW")
    N, C, H, W = input.shape
    outH = (H + 2 * padding[0] - kernel[0]) // stride[0] + 1
    outW = (W + 2 * padding[1] - kernel[1]) // stride[1] + 1
    if outH < 1 or outW < 1:
        raise ValueError("maxpool2d: kernel larger than the padded input")
    pooled = torch.empty((N, C, outH, outW), dtype=input.dtype, device=input.device)
    idx = torch.empty((N * C * outH * outW,), dtype=torch.int64, device=input.device)
    with torch.cuda.device(input.device):
        _capi.check(getattr(_capi.lib(), f"am_maxpool2d_forward_{suf}")(
            _stream_ptr(input), N, C, H, W, kernel[0], kernel[1], padding[0], padding[1], stride[0], stride[1],
            input.data_ptr(), pooled.data_ptr(), idx.data_ptr()))
    return idx, pooled


def maxpool2d_backward(cached_input_shape, cached_max_indices: torch.Tensor, grad_output: torch.Tensor,
                       windows_overlap: bool = True, relu_cached: torch.Tensor | None = None) -> torch.Tensor:
    """gradInput = zeros; gradInput[max_indices[i]] = gradOutput[i] (nnp_maxpooling.nim:68-83).  Pass
    windows_overlap=False when stride >= kernel (every input belongs to one window): skips the pass that
    reproduces the serial reference's "last writer wins".
    relu_cached (the tensor the pooled layer's input was relu'd from / to, same shape as the result): additionally applies
    relu_backward (nnp_activation.nim:65-70) in the same pass — identical bits to the two separate calls."""
    suf = _fcheck("maxpool2d_backward", grad_output)
    if cached_max_indices.dtype != torch.int64 or not cached_max_indices.is_cuda or not cached_max_indices.is_contiguous():
        raise TypeError("maxpool2d_backward: max_indices must be a contiguous int64 GPU tensor")
    if cached_max_indices.numel() != grad_output.numel():
        raise IndexError("maxpool2d_backward: gradOutput and max_indices sizes differ")
    gin = torch.empty(tuple(cached_input_shape), dtype=grad_output.dtype, device=grad_output.device)
    if relu_cached is not None:
        if (relu_cached.dtype != grad_output.dtype or tuple(relu_cached.shape) != tuple(cached_input_shape)
                or not relu_cached.is_cuda or not relu_cached.is_contiguous()):
            raise ValueError("maxpool2d_backward: relu_cached must be a contiguous GPU tensor of the input's shape and dtype")
    with torch.cuda.device(grad_output.device):
        if relu_cached is None:
            _capi.check(getattr(_capi.lib(), f"am_maxpool2d_backward_{suf}")(
                _stream_ptr(grad_output), gin.numel(), grad_output.numel(), cached_max_indices.data_ptr(),
                grad_output.data_ptr(), gin.data_ptr(), 1 if windows_overlap else 0))
        else:
            _capi.check(getattr(_capi.lib(), f"am_maxpool2d_backward_relu_{suf}")(
                _stream_ptr(grad_output), gin.numel(), grad_output.numel(), cached_max_indices.data_ptr(),
                grad_output.data_ptr(), relu_cached.data_ptr(), gin.data_ptr(), 1 if windows_overlap else 0))
    return gin


def linear(input: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """Y = x * W^T (+ b): input [batch, in], weight [out, in], bias [1, out] (nnp_linear.nim:20-37)."""
    suf = _fcheck("linear", input, weight, bias)
    if input.dim() != 2 or weight.dim() != 2 or input.shape[1] != weight.shape[1]:
        raise IndexError("linear: input [batch, in] and weight [out, in] do not match")
    if bias is not None and bias.numel() != weight.shape[0]:
        raise IndexError("linear: bias must have out_features elements")
    out = torch.empty((input.shape[0], weight.shape[0]), dtype=input.dtype, device=input.device)
    with torch.cuda.device(input.device):
        _capi.check(getattr(_capi.lib(), f"am_linear_forward_{suf}")(
            _stream_ptr(input), input.shape[0], input.shape[1], weight.shape[0], input.data_ptr(), weight.data_ptr(),
            bias.data_ptr() if bias is not None else None, out.data_ptr()))
    return out


def linear_backward(input: torch.Tensor, weight: torch.Tensor, grad_output: torch.Tensor, with_bias: bool = True):
    """-> (gradInput = gO*W, gradWeight = gO^T*x, gradBias = sum(gO, axis 0) [1, out] or None) (nnp_linear.nim:39-66)."""
    suf = _fcheck("linear_backward", input, weight, grad_output)
    if grad_output.shape != (input.shape[0], weight.shape[0]) or input.shape[1] != weight.shape[1]:
        raise IndexError("linear_backward: shapes do not match")
    gi = torch.empty_like(input)
    gw = torch.empty_like(weight)
    gb = torch.empty((1, weight.shape[0]), dtype=input.dtype, device=input.device) if with_bias else None
    with torch.cuda.device(input.device):
        _capi.check(getattr(_capi.lib(), f"am_linear_backward_{suf}")(
            _stream_ptr(input), input.shape[0], input.shape[1], weight.shape[0], input.data_ptr(), weight.data_ptr(),
            grad_output.data_ptr(), gi.data_ptr(), gw.data_ptr(), gb.data_ptr() if gb is not None else None))
    return gi, gw, gb


def _labels(target: torch.Tensor, batch: int, device):
    if target.numel() != batch:
        raise IndexError("sparse_softmax_cross_entropy: target must hold one label per sample")
    return target.to(device=device, dtype=torch.int64).contiguous().reshape(-1)


def sparse_softmax_cross_entropy_dev(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Loss as a 0-dim device tensor (no host synchronisation)."""
    if input.dim() != 2:
        raise ValueError("sparse_softmax_cross_entropy: input must be [batch, features]")
    suf = _fcheck("sparse_softmax_cross_entropy", input.contiguous())
    if not input.is_cuda:
        raise ValueError("sparse_softmax_cross_entropy: input must live on the GPU")
    lab = _labels(target, input.shape[0], input.device)
    loss = torch.empty((), dtype=input.dtype, device=input.device)
    with torch.cuda.device(input.device):
        _capi.check(getattr(_capi.lib(), f"am_sparse_softmax_cross_entropy_{suf}")(
            _stream_ptr(input), input.shape[0], input.shape[1], input.data_ptr(), input.stride(0), input.stride(1),
            lab.data_ptr(), loss.data_ptr()))
    return loss


def sparse_softmax_cross_entropy(input: torch.Tensor, target: torch.Tensor) -> float:
    """Mean over the batch of logsumexp(x_i) - x_i[label_i] (nnp_softmax_cross_entropy.nim:100-178); returns the
    scalar like the reference does (one device-to-host read)."""
    return float(sparse_softmax_cross_entropy_dev(input, target).item())


def sparse_softmax_cross_entropy_backward(gradient, cached_tensor: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """gradient * (softmax(cached) - onehot(target)) / batch (nnp_softmax_cross_entropy.nim:219-252); `gradient` is a
    scalar or a one-element tensor."""
    if cached_tensor.dim() != 2 or not cached_tensor.is_cuda:
        raise ValueError("sparse_softmax_cross_entropy_backward: cached tensor must be a [batch, features] GPU tensor")
    if cached_tensor.dtype not in _FSUF:
        raise TypeError("sparse_softmax_cross_entropy_backward: float32 or float64")
    suf = _FSUF[cached_tensor.dtype]
    g = float(gradient.reshape(-1)[0].item()) if isinstance(gradient, torch.Tensor) else float(gradient)
    lab = _labels(target, cached_tensor.shape[0], cached_tensor.device)
    out = torch.empty(tuple(cached_tensor.shape), dtype=cached_tensor.dtype, device=cached_tensor.device)
    with torch.cuda.device(cached_tensor.device):
        _capi.check(getattr(_capi.lib(), f"am_sparse_softmax_cross_entropy_backward_{suf}")(
            _stream_ptr(cached_tensor), cached_tensor.shape[0], cached_tensor.shape[1], g, cached_tensor.data_ptr(),
            cached_tensor.stride(0), cached_tensor.stride(1), lab.data_ptr(), out.data_ptr()))
    return out
