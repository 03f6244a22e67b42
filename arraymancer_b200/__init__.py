"""arraymancer_b200 — B200-native (sm_100a) drop-in for Arraymancer's dense-contraction hot path:
`gemm_strided` / `CudaTensor *` for float32, float64, int32, int64 and the im2col+GEMM conv2d
forward/backward (as a fused implicit GEMM).  The arithmetic lives in
libarraymancer_b200.so (C ABI: include/am_b200.h); this package is the thin host-side mirror of
the reference's operator interface.  There is no CPU fallback."""
from . import _capi
from ._capi import (AmError, F32_AUTO, F32_SIMT, F32_TC, F32_TC_1CTA, F64_AUTO, F64_DMMA, F64_SIMT, set_f32_path,
                    set_f64_path, version)
from .cuda_tensor import CudaTensor, PackedF32, cuda, cublas_gemm, gemm, gemm_packed, gemm_packed_bcast, gemm_strided, gemm_strided_batched, matmul
from .nn_primitives import (conv2d, conv2d_backward, conv_out_dims, linear, linear_backward, maxpool2d, maxpool2d_backward, relu,
                            relu_backward, sparse_softmax_cross_entropy, sparse_softmax_cross_entropy_backward,
                            sparse_softmax_cross_entropy_dev)

__all__ = ["AmError", "CudaTensor", "cuda", "cublas_gemm", "gemm", "gemm_strided", "gemm_strided_batched", "gemm_packed", "gemm_packed_bcast", "PackedF32", "matmul", "conv2d",
           "conv2d_backward", "conv_out_dims", "relu", "relu_backward", "maxpool2d", "maxpool2d_backward", "linear", "linear_backward",
           "sparse_softmax_cross_entropy", "sparse_softmax_cross_entropy_dev", "sparse_softmax_cross_entropy_backward", "set_f32_path", "set_f64_path", "version", "F64_AUTO", "F64_SIMT", "F64_DMMA", "F32_AUTO", "F32_SIMT", "F32_TC",
           "F32_TC_1CTA"]
