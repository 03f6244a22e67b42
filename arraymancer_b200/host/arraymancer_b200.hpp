// arraymancer_b200.hpp — C++ host-side mirror of the reference's CudaTensor operator interface for the
// dense-contraction path, layered on the C ABI of libarraymancer_b200.so (include/am_b200.h).
//
// The reference's host language is Nim (not available in this image); this header stands where the Nim module of
// INTEGRATION.md would, with the same names, argument meaning and error behaviour as
//   tensor/data_structure.nim:44-58            CudaTensor[T] (shape, strides, offset, ref-counted device storage)
//   tensor/init_cuda.nim:23-59                 cuda() (column-major, H2D) / cpu() (blocking D2H)
//   tensor/operators_blas_l2l3_cuda.nim:43-87  `*`, cudaMM_C_eq_aAB_p_bC
//   nn_primitives/nnp_conv2d_cudnn.nim:20-204  conv2d / conv2d_backward (SizeHW = array[2, int])
// (paths relative to /root/reference/src/arraymancer/).  ValueError -> std::invalid_argument,
// IndexDefect -> std::out_of_range, CUDA status -> std::runtime_error via am_last_error().
#pragma once
#include <cuda_runtime.h>

#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/am_b200.h"

namespace arraymancer_b200 {

inline void amCheck(int status) {   // nimcuda-style `check`
  if (status != AM_OK) throw std::runtime_error(std::string("arraymancer_b200: ") + am_last_error());
}
inline void cudaCheck(cudaError_t e) {
  if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e));
}

using SizeHW = std::array<int64_t, 2>;   // backend/cudnn_conv_interface.nim:28

template <class T>
struct CudaStorage {   // CudaStorage[T]: cudaMalloc'd block freed when the last tensor referencing it dies
  T* data = nullptr;
  size_t len = 0;
  explicit CudaStorage(size_t n) : len(n) { cudaCheck(cudaMalloc((void**)&data, (n ? n : 1) * sizeof(T))); }
  ~CudaStorage() { if (data) cudaFree(data); }
  CudaStorage(const CudaStorage&) = delete;
};

template <class T>
struct CudaTensor {
  std::vector<int64_t> shape, strides;   // strides in elements
  int64_t offset = 0;
  std::shared_ptr<CudaStorage<T>> storage;

  int rank() const { return (int)shape.size(); }
  int64_t size() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
  T* get_offset_ptr() const { return storage->data + offset; }   // data_structure.nim:193-198

  // newCudaTensor (p_init_cuda.nim:19-46): uninitialised, COLUMN-major by default like the reference
  static CudaTensor make(const std::vector<int64_t>& shp, bool colMajor = true) {
    CudaTensor t;
    t.shape = shp;
    t.strides.assign(shp.size(), 1);
    int64_t acc = 1;
    if (colMajor) { for (size_t i = 0; i < shp.size(); i++) { t.strides[i] = acc; acc *= shp[i]; } }
    else { for (size_t i = shp.size(); i-- > 0;) { t.strides[i] = acc; acc *= shp[i]; } }
    t.storage = std::make_shared<CudaStorage<T>>((size_t)acc);
    return t;
  }
  CudaTensor transpose() const {          // stride swap, no copy
    CudaTensor t = *this;
    std::swap(t.shape[0], t.shape[1]);
    std::swap(t.strides[0], t.strides[1]);
    return t;
  }
  // basic slicing with a step per axis (negative steps allowed): offset += start*stride; stride *= step
  CudaTensor slice(int axis, int64_t start, int64_t count, int64_t step) const {
    CudaTensor t = *this;
    t.offset += start * strides[axis];
    t.shape[axis] = count;
    t.strides[axis] = strides[axis] * step;
    return t;
  }
};

// `t.cuda()` (init_cuda.nim:23-41): host row-major data -> column-major device tensor
template <class T>
CudaTensor<T> cuda(const std::vector<T>& rowMajor, const std::vector<int64_t>& shape) {
  CudaTensor<T> t = CudaTensor<T>::make(shape, /*colMajor=*/true);
  const int64_t n = t.size();
  if ((int64_t)rowMajor.size() != n) throw std::invalid_argument("cuda(): data size does not match the shape");
  std::vector<T> cm((size_t)n);
  if (shape.size() == 2) {
    for (int64_t r = 0; r < shape[0]; r++)
      for (int64_t c = 0; c < shape[1]; c++) cm[(size_t)(r + c * shape[0])] = rowMajor[(size_t)(r * shape[1] + c)];
  } else {
    cm = rowMajor;   // rank-1 (and NCHW buffers handed over as-is via make(..., false))
  }
  cudaCheck(cudaMemcpy(t.storage->data, cm.data(), (size_t)n * sizeof(T), cudaMemcpyHostToDevice));
  return t;
}
// `.cpu()` (init_cuda.nim:43-59): blocking D2H; returned in logical row-major order
template <class T>
std::vector<T> cpu(const CudaTensor<T>& t) {
  std::vector<T> raw(t.storage->len);
  cudaCheck(cudaMemcpy(raw.data(), t.storage->data, raw.size() * sizeof(T), cudaMemcpyDeviceToHost));
  std::vector<T> out((size_t)t.size());
  if (t.rank() == 2) {
    for (int64_t r = 0; r < t.shape[0]; r++)
      for (int64_t c = 0; c < t.shape[1]; c++)
        out[(size_t)(r * t.shape[1] + c)] = raw[(size_t)(t.offset + r * t.strides[0] + c * t.strides[1])];
  } else {
    for (int64_t i = 0; i < t.size(); i++) out[(size_t)i] = raw[(size_t)(t.offset + i * (t.rank() ? t.strides.back() : 1))];
  }
  return out;
}

namespace detail {
template <class T> struct Abi;
#define AMB200_ABI(T, SUF)                                                                                   \
  template <> struct Abi<T> {                                                                                \
    static int gemm(am_stream_t s, int64_t M, int64_t N, int64_t K, T al, const T* A, int64_t ra, int64_t ca, \
                    const T* B, int64_t rb, int64_t cb, T be, T* C, int64_t rc, int64_t cc) {                 \
      return am_gemm_strided_##SUF(s, M, N, K, al, A, ra, ca, B, rb, cb, be, C, rc, cc);                      \
    }                                                                                                        \
    static int conv_fwd(am_stream_t s, const am_conv2d_desc* d, const T* x, const T* k, const T* b, T* y) {  \
      return am_conv2d_forward_##SUF(s, d, x, k, b, y);                                                       \
    }                                                                                                        \
    static int conv_bwd(am_stream_t s, const am_conv2d_desc* d, const T* x, const T* k, const T* go, T* gi,  \
                        T* gk, T* gb) {                                                                       \
      return am_conv2d_backward_##SUF(s, d, x, k, go, gi, gk, gb);                                            \
    }                                                                                                        \
  };
AMB200_ABI(float, f32)
AMB200_ABI(double, f64)
AMB200_ABI(int32_t, i32)
AMB200_ABI(int64_t, i64)
#undef AMB200_ABI
}  // namespace detail

// gemm(alpha, A, B, beta, C) — cudaMM_C_eq_aAB_p_bC (operators_blas_l2l3_cuda.nim:43-72); any strided views
template <class T>
void gemm(T alpha, const CudaTensor<T>& a, const CudaTensor<T>& b, T beta, CudaTensor<T>& c, cudaStream_t st = nullptr) {
  if (a.rank() != 2 || b.rank() != 2 || c.rank() != 2) throw std::invalid_argument("gemm: inputs must be matrices");
  if (a.shape[1] != b.shape[0] || c.shape[0] != a.shape[0] || c.shape[1] != b.shape[1])
    throw std::out_of_range("gemm: shape mismatch");                  // check_matmat -> IndexDefect
  amCheck(detail::Abi<T>::gemm((am_stream_t)st, a.shape[0], b.shape[1], a.shape[1], alpha, a.get_offset_ptr(), a.strides[0],
                               a.strides[1], b.get_offset_ptr(), b.strides[0], b.strides[1], beta, c.get_offset_ptr(),
                               c.strides[0], c.strides[1]));
}

// `a * b` (operators_blas_l2l3_cuda.nim:74-87): matrix x matrix -> fresh column-major result, alpha = 1, beta = 0
template <class T>
CudaTensor<T> operator*(const CudaTensor<T>& a, const CudaTensor<T>& b) {
  if (a.rank() == 2 && b.rank() == 2) {
    if (a.shape[1] != b.shape[0]) throw std::out_of_range("matmul: inner dimensions differ");
    CudaTensor<T> r = CudaTensor<T>::make({a.shape[0], b.shape[1]});
    gemm<T>(T(1), a, b, T(0), r);
    return r;
  }
  throw std::invalid_argument("Matrix-Matrix or Matrix-Vector multiplication valid only if first Tensor is a Matrix and "
                              "second is a Matrix or Vector");
}

inline am_conv2d_desc convDesc(const std::vector<int64_t>& in, const std::vector<int64_t>& k, SizeHW pad, SizeHW st, SizeHW dil) {
  if (in.size() != 4 || k.size() != 4) throw std::invalid_argument("conv2d: input and kernel must be rank-4");
  if (in[1] != k[1]) throw std::out_of_range("conv2d: channel mismatch");
  return am_conv2d_desc{in[0], in[1], in[2], in[3], k[0], k[2], k[3], pad[0], pad[1], st[0], st[1], dil[0], dil[1]};
}

// conv2d(input, kernel, bias, padding, strides, dilation) — nnp_conv2d_cudnn.nim:20-72; tensors are C-contiguous NCHW;
// bias [Cout,1,1] or a rank-0 (empty) tensor for "no bias"
template <class T>
CudaTensor<T> conv2d(const CudaTensor<T>& input, const CudaTensor<T>& kernel, const CudaTensor<T>& bias,
                     SizeHW padding = {0, 0}, SizeHW strides = {1, 1}, SizeHW dilation = {1, 1}, cudaStream_t st = nullptr) {
  am_conv2d_desc d = convDesc(input.shape, kernel.shape, padding, strides, dilation);
  int64_t ho = 0, wo = 0;
  amCheck(am_conv2d_out_dims(&d, &ho, &wo));
  if (ho <= 0 || wo <= 0) throw std::invalid_argument("conv2d: kernel larger than the padded input");
  CudaTensor<T> out = CudaTensor<T>::make({d.N, d.Cout, ho, wo}, /*colMajor=*/false);
  amCheck(detail::Abi<T>::conv_fwd((am_stream_t)st, &d, input.get_offset_ptr(), kernel.get_offset_ptr(),
                                   bias.rank() > 0 ? bias.get_offset_ptr() : nullptr, out.get_offset_ptr()));
  return out;
}

// conv2d_backward(..., grad_output, grad_input, grad_kernel, grad_bias) — nnp_conv2d_cudnn.nim:74-204
template <class T>
void conv2d_backward(const CudaTensor<T>& input, const CudaTensor<T>& kernel, const CudaTensor<T>& bias, SizeHW padding,
                     SizeHW strides, SizeHW dilation, const CudaTensor<T>& grad_output, CudaTensor<T>& grad_input,
                     CudaTensor<T>& grad_kernel, CudaTensor<T>& grad_bias, cudaStream_t st = nullptr) {
  am_conv2d_desc d = convDesc(input.shape, kernel.shape, padding, strides, dilation);
  grad_input = CudaTensor<T>::make(input.shape, false);
  grad_kernel = CudaTensor<T>::make(kernel.shape, false);
  const bool has_bias = bias.rank() > 0;
  if (has_bias) grad_bias = CudaTensor<T>::make(bias.shape, false);
  amCheck(detail::Abi<T>::conv_bwd((am_stream_t)st, &d, input.get_offset_ptr(), kernel.get_offset_ptr(),
                                   grad_output.get_offset_ptr(), grad_input.get_offset_ptr(), grad_kernel.get_offset_ptr(),
                                   has_bias ? grad_bias.get_offset_ptr() : nullptr));
}


// ---- the operators around the contractions, float32 / float64, C-contiguous tensors
namespace detail {
template <class T> struct NnAbi;
#define AMB200_NN(T, SUF)                                                                                              \
  template <> struct NnAbi<T> {                                                                                        \
    static int relu(am_stream_t s, int64_t n, const T* x, T* y) { return am_relu_forward_##SUF(s, n, x, y); }          \
    static int relu_bwd(am_stream_t s, int64_t n, const T* g, const T* c, T* o) { return am_relu_backward_##SUF(s, n, g, c, o); } \
    static int pool(am_stream_t s, int64_t N, int64_t C, int64_t H, int64_t W, int64_t kH, int64_t kW, int64_t pH, int64_t pW, \
                    int64_t sH, int64_t sW, const T* x, T* y, int64_t* idx) {                                          \
      return am_maxpool2d_forward_##SUF(s, N, C, H, W, kH, kW, pH, pW, sH, sW, x, y, idx);                             \
    }                                                                                                                  \
    static int pool_bwd(am_stream_t s, int64_t nin, int64_t nout, const int64_t* idx, const T* go, T* gi, int ov) {    \
      return am_maxpool2d_backward_##SUF(s, nin, nout, idx, go, gi, ov);                                               \
    }                                                                                                                  \
    static int linear(am_stream_t s, int64_t b, int64_t in, int64_t out, const T* x, const T* w, const T* bias, T* y) { \
      return am_linear_forward_##SUF(s, b, in, out, x, w, bias, y);                                                    \
    }                                                                                                                  \
    static int linear_bwd(am_stream_t s, int64_t b, int64_t in, int64_t out, const T* x, const T* w, const T* go, T* gi, \
                          T* gw, T* gb) {                                                                              \
      return am_linear_backward_##SUF(s, b, in, out, x, w, go, gi, gw, gb);                                            \
    }                                                                                                                  \
    static int ssce(am_stream_t s, int64_t b, int64_t f, const T* x, int64_t rs, int64_t cs, const int64_t* lab, T* loss) { \
      return am_sparse_softmax_cross_entropy_##SUF(s, b, f, x, rs, cs, lab, loss);                                     \
    }                                                                                                                  \
    static int ssce_bwd(am_stream_t s, int64_t b, int64_t f, T g, const T* x, int64_t rs, int64_t cs, const int64_t* lab, \
                        T* out) {                                                                                      \
      return am_sparse_softmax_cross_entropy_backward_##SUF(s, b, f, g, x, rs, cs, lab, out);                          \
    }                                                                                                                  \
  };
AMB200_NN(float, f32)
AMB200_NN(double, f64)
#undef AMB200_NN
}  // namespace detail

// relu / relu_backward — nnp_activation.nim:35-36, 65-70
template <class T>
CudaTensor<T> relu(const CudaTensor<T>& t, cudaStream_t st = nullptr) {
  CudaTensor<T> out = CudaTensor<T>::make(t.shape, false);
  amCheck(detail::NnAbi<T>::relu((am_stream_t)st, t.size(), t.get_offset_ptr(), out.get_offset_ptr()));
  return out;
}
template <class T>
CudaTensor<T> relu_backward(const CudaTensor<T>& gradient, const CudaTensor<T>& cached, cudaStream_t st = nullptr) {
  if (gradient.shape != cached.shape) throw std::out_of_range("relu_backward: shapes differ");
  CudaTensor<T> out = CudaTensor<T>::make(gradient.shape, false);
  amCheck(detail::NnAbi<T>::relu_bwd((am_stream_t)st, gradient.size(), gradient.get_offset_ptr(), cached.get_offset_ptr(),
                                     out.get_offset_ptr()));
  return out;
}

// maxpool2d(input, kernel, padding, stride) -> (max_indices, maxpooled) — nnp_maxpooling.nim:19-66
template <class T>
struct MaxPoolResult { CudaTensor<int64_t> max_indices; CudaTensor<T> maxpooled; };
template <class T>
MaxPoolResult<T> maxpool2d(const CudaTensor<T>& input, SizeHW kernel, SizeHW padding = {0, 0}, SizeHW stride = {1, 1},
                           cudaStream_t st = nullptr) {
  if (input.rank() != 4) throw std::invalid_argument("maxpool2d: input must be rank-4 NCHW");
  const int64_t N = input.shape[0], C = input.shape[1], H = input.shape[2], W = input.shape[3];
  const int64_t oh = (H + 2 * padding[0] - kernel[0]) / stride[0] + 1, ow = (W + 2 * padding[1] - kernel[1]) / stride[1] + 1;
  if (oh < 1 || ow < 1) throw std::invalid_argument("maxpool2d: kernel larger than the padded input");
  MaxPoolResult<T> r;
  r.max_indices = CudaTensor<int64_t>::make({N * C * oh * ow}, false);
  r.maxpooled = CudaTensor<T>::make({N, C, oh, ow}, false);
  amCheck(detail::NnAbi<T>::pool((am_stream_t)st, N, C, H, W, kernel[0], kernel[1], padding[0], padding[1], stride[0], stride[1],
                                 input.get_offset_ptr(), r.maxpooled.get_offset_ptr(), r.max_indices.get_offset_ptr()));
  return r;
}
// maxpool2d_backward(cached_input_shape, cached_max_indices, gradOutput) — nnp_maxpooling.nim:68-83
template <class T>
CudaTensor<T> maxpool2d_backward(const std::vector<int64_t>& input_shape, const CudaTensor<int64_t>& max_indices,
                                 const CudaTensor<T>& grad_output, bool windows_overlap = true, cudaStream_t st = nullptr) {
  if (max_indices.size() != grad_output.size()) throw std::out_of_range("maxpool2d_backward: sizes differ");
  CudaTensor<T> gi = CudaTensor<T>::make(input_shape, false);
  amCheck(detail::NnAbi<T>::pool_bwd((am_stream_t)st, gi.size(), grad_output.size(), max_indices.get_offset_ptr(),
                                     grad_output.get_offset_ptr(), gi.get_offset_ptr(), windows_overlap ? 1 : 0));
  return gi;
}

// linear(input [batch,in], weight [out,in], bias [1,out] or rank-0) -> [batch,out] — nnp_linear.nim:20-37
template <class T>
CudaTensor<T> linear(const CudaTensor<T>& input, const CudaTensor<T>& weight, const CudaTensor<T>& bias, cudaStream_t st = nullptr) {
  if (input.rank() != 2 || weight.rank() != 2 || input.shape[1] != weight.shape[1]) throw std::out_of_range("linear: shapes do not match");
  CudaTensor<T> out = CudaTensor<T>::make({input.shape[0], weight.shape[0]}, false);
  amCheck(detail::NnAbi<T>::linear((am_stream_t)st, input.shape[0], input.shape[1], weight.shape[0], input.get_offset_ptr(),
                                   weight.get_offset_ptr(), bias.rank() > 0 ? bias.get_offset_ptr() : nullptr, out.get_offset_ptr()));
  return out;
}
// linear_backward(input, weight, gradOutput, gradInput, gradWeight, gradBias) — nnp_linear.nim:39-66
template <class T>
void linear_backward(const CudaTensor<T>& input, const CudaTensor<T>& weight, const CudaTensor<T>& grad_output,
                     CudaTensor<T>& grad_input, CudaTensor<T>& grad_weight, CudaTensor<T>* grad_bias, cudaStream_t st = nullptr) {
  grad_input = CudaTensor<T>::make(input.shape, false);
  grad_weight = CudaTensor<T>::make(weight.shape, false);
  if (grad_bias) *grad_bias = CudaTensor<T>::make({1, weight.shape[0]}, false);
  amCheck(detail::NnAbi<T>::linear_bwd((am_stream_t)st, input.shape[0], input.shape[1], weight.shape[0], input.get_offset_ptr(),
                                       weight.get_offset_ptr(), grad_output.get_offset_ptr(), grad_input.get_offset_ptr(),
                                       grad_weight.get_offset_ptr(), grad_bias ? grad_bias->get_offset_ptr() : nullptr));
}

// sparse_softmax_cross_entropy(input [batch,features], target [batch]) -> scalar — nnp_softmax_cross_entropy.nim:100-178
template <class T>
T sparse_softmax_cross_entropy(const CudaTensor<T>& input, const CudaTensor<int64_t>& target, cudaStream_t st = nullptr) {
  if (input.rank() != 2 || target.size() != input.shape[0]) throw std::out_of_range("sparse_softmax_cross_entropy: shapes do not match");
  CudaTensor<T> loss = CudaTensor<T>::make({1}, false);
  amCheck(detail::NnAbi<T>::ssce((am_stream_t)st, input.shape[0], input.shape[1], input.get_offset_ptr(), input.strides[0],
                                 input.strides[1], target.get_offset_ptr(), loss.get_offset_ptr()));
  T h;
  cudaCheck(cudaMemcpyAsync(&h, loss.get_offset_ptr(), sizeof(T), cudaMemcpyDeviceToHost, st));
  cudaCheck(cudaStreamSynchronize(st));
  return h;
}
// sparse_softmax_cross_entropy_backward(gradient, cached, target) — nnp_softmax_cross_entropy.nim:219-252
template <class T>
CudaTensor<T> sparse_softmax_cross_entropy_backward(T gradient, const CudaTensor<T>& cached, const CudaTensor<int64_t>& target,
                                                    cudaStream_t st = nullptr) {
  CudaTensor<T> out = CudaTensor<T>::make(cached.shape, false);
  amCheck(detail::NnAbi<T>::ssce_bwd((am_stream_t)st, cached.shape[0], cached.shape[1], gradient, cached.get_offset_ptr(),
                                     cached.strides[0], cached.strides[1], target.get_offset_ptr(), out.get_offset_ptr()));
  return out;
}

}  // namespace arraymancer_b200
