// Boundary extensions of the C ABI (include/am_b200.h, second half):
//   * am_gemm_strided_batched_*  — `batch` independent products in one call (tensor/backend/cublas.nim:172-208
//                                  `cublas_gemmStridedBatched`; the reference's own "TODO: batch matmul" over images,
//                                  nn_primitives/fallback/conv.nim:99)
//   * am_conv2d_*_strided_*      — conv entries that take 4 element strides per tensor, the way the reference builds its
//                                  descriptors from `t.strides[0..3]` (nn_primitives/backend/cudnn.nim:59-75): the Nim
//                                  `conv2d_backward` no longer needs the `asContiguous` that is `{.error.}` on CUDA
//                                  (nnp_conv2d_cudnn.nim:99-101)
//   * am_packed_floats_f32 / am_pack_f32_*_into / am_packed_wrap_f32 — packed operands in caller-owned memory (K slices
//                                  exchanged between GPUs)
#include <cstring>

#include "am_common.cuh"
#include "gemm_dispatch.h"

namespace am {

// ------------------------------------------------------------------ 4-D strided <-> dense copies
struct Dims4 { int64_t n[4]; int64_t s[4]; };

// dense (C-contiguous) index i <-> strided element; one of the two sides is always dense, so that side is coalesced
template <class T, bool ToDense>
__global__ void __launch_bounds__(256) strided_copy4_kernel(const T* __restrict__ src, T* __restrict__ dst, Dims4 d, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int64_t i3 = r % d.n[3]; r /= d.n[3];
    const int64_t i2 = r % d.n[2]; r /= d.n[2];
    const int64_t i1 = r % d.n[1]; r /= d.n[1];
    const int64_t off = r * d.s[0] + i1 * d.s[1] + i2 * d.s[2] + i3 * d.s[3];
    if (ToDense) dst[i] = src[off];
    else dst[off] = src[i];
  }
}

static bool is_dense4(const int64_t n[4], const int64_t* s) {
  if (!s) return true;
  int64_t acc = 1;
  for (int d = 3; d >= 0; d--) {
    if (n[d] != 1 && s[d] != acc) return false;
    acc *= n[d];
  }
  return true;
}

// A tensor argument of the strided conv entries: if its strides are not the dense NCHW ones, `dense` is a workspace copy.
template <class T>
struct StridedArg {
  const T* user = nullptr; T* dense = nullptr; Dims4 d{}; int64_t total = 0; bool copied = false;
  int init(int slot, const T* p, const int64_t n[4], const int64_t* s) {
    user = p; total = n[0] * n[1] * n[2] * n[3];
    for (int i = 0; i < 4; i++) { d.n[i] = n[i]; d.s[i] = s ? s[i] : 0; }
    if (!p || total == 0 || is_dense4(n, s)) { dense = const_cast<T*>(p); copied = false; return AM_OK; }
    void* w = nullptr;
    int rc = workspace(slot, (size_t)total * sizeof(T), &w);
    if (rc) return rc;
    dense = (T*)w; copied = true;
    return AM_OK;
  }
  int gather(cudaStream_t st) const {       // user (strided) -> dense workspace
    if (!copied) return AM_OK;
    const unsigned blocks = (unsigned)(ceil_div(total, 256) < 8 * (int64_t)sm_count() ? ceil_div(total, 256) : 8 * (int64_t)sm_count());
    strided_copy4_kernel<T, true><<<blocks, 256, 0, st>>>(user, dense, d, total);
    g_launch_count++;
    AM_CUDA_TRY(cudaGetLastError());
    return AM_OK;
  }
  int scatter(cudaStream_t st) const {      // dense workspace -> user (strided)
    if (!copied) return AM_OK;
    const unsigned blocks = (unsigned)(ceil_div(total, 256) < 8 * (int64_t)sm_count() ? ceil_div(total, 256) : 8 * (int64_t)sm_count());
    strided_copy4_kernel<T, false><<<blocks, 256, 0, st>>>(dense, const_cast<T*>(user), d, total);
    g_launch_count++;
    AM_CUDA_TRY(cudaGetLastError());
    return AM_OK;
  }
};

template <class T>
static int conv_forward_strided(cudaStream_t st, const am_conv2d_desc& d, const T* in, const int64_t* is, const T* k,
                                const int64_t* ks, const T* bias, int64_t bias_stride, T* out, const int64_t* os, int act) {
  int64_t Ho = 0, Wo = 0;
  int rc = am_conv2d_out_dims(&d, &Ho, &Wo);
  if (rc) return rc;
  const int64_t nin[4] = {d.N, d.C, d.H, d.W}, nk[4] = {d.Cout, d.C, d.kH, d.kW}, nout[4] = {d.N, d.Cout, Ho, Wo};
  const int64_t nb[4] = {1, 1, 1, d.Cout}, sb[4] = {0, 0, 0, bias_stride};
  StridedArg<T> ai, ak, ab, ao;
  if ((rc = ai.init(kWsStrideIn, in, nin, is)) || (rc = ak.init(kWsStrideK, k, nk, ks)) ||
      (rc = ab.init(kWsStrideGk, bias, nb, bias && bias_stride != 1 ? sb : nullptr)) || (rc = ao.init(kWsStrideOut, out, nout, os)))
    return rc;
  if ((rc = ai.gather(st)) || (rc = ak.gather(st)) || (rc = ab.gather(st))) return rc;
  if ((rc = conv2d_forward<T>(st, d, ai.dense, ak.dense, ab.dense, ao.dense, act))) return rc;
  return ao.scatter(st);
}

template <class T>
static int conv_backward_strided(cudaStream_t st, const am_conv2d_desc& d, const T* in, const int64_t* is, const T* k,
                                 const int64_t* ks, const T* go, const int64_t* gos, T* gi, const int64_t* gis, T* gk,
                                 const int64_t* gks, T* gb, int64_t gb_stride) {
  int64_t Ho = 0, Wo = 0;
  int rc = am_conv2d_out_dims(&d, &Ho, &Wo);
  if (rc) return rc;
  const int64_t nin[4] = {d.N, d.C, d.H, d.W}, nk[4] = {d.Cout, d.C, d.kH, d.kW}, nout[4] = {d.N, d.Cout, Ho, Wo};
  const int64_t nb[4] = {1, 1, 1, d.Cout}, sb[4] = {0, 0, 0, gb_stride};
  StridedArg<T> ai, ak, ago, agi, agk, agb;
  // grad_input shares the slot family of the forward output; grad_bias is tiny and gets the NN slot
  if ((rc = ai.init(kWsStrideIn, in, nin, is)) || (rc = ak.init(kWsStrideK, k, nk, ks)) || (rc = ago.init(kWsStrideGo, go, nout, gos)) ||
      (rc = agi.init(kWsStrideOut, gi, nin, gis)) || (rc = agk.init(kWsStrideGk, gk, nk, gks)))
    return rc;
  // grad_bias with a stride (e.g. a [Cout,1,1] view of a larger buffer): computed densely into a small scratch
  T* gb_dense = gb;
  void* gbw = nullptr;
  if (gb && gb_stride != 1) {
    if ((rc = workspace(kWsNn, (size_t)d.Cout * sizeof(T), &gbw))) return rc;
    gb_dense = (T*)gbw;
  }
  if ((rc = ai.gather(st)) || (rc = ak.gather(st)) || (rc = ago.gather(st))) return rc;
  if ((rc = conv2d_backward<T>(st, d, ai.dense, ak.dense, ago.dense, agi.dense, agk.dense, gb_dense))) return rc;
  if ((rc = agi.scatter(st)) || (rc = agk.scatter(st))) return rc;
  if (gb && gb_stride != 1) {
    agb.user = gb; agb.dense = gb_dense; agb.total = d.Cout; agb.copied = true;
    for (int i = 0; i < 4; i++) { agb.d.n[i] = nb[i]; agb.d.s[i] = sb[i]; }
    if ((rc = agb.scatter(st))) return rc;
  }
  return AM_OK;
}

// ------------------------------------------------------------------ batched GEMM
template <class T>
static int gemm_batched(cudaStream_t st, int64_t batch, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t rsA,
                        int64_t csA, int64_t bsA, const T* B, int64_t rsB, int64_t csB, int64_t bsB, T beta, T* C, int64_t rsC,
                        int64_t csC, int64_t bsC) {
  if (batch < 0 || M < 0 || N < 0 || K < 0) { set_last_error("gemm_strided_batched: negative dimension"); return AM_ERR_INVALID; }
  if (batch == 0 || M == 0 || N == 0 || K == 0) return AM_OK;      // K == 0 leaves C untouched, like am_gemm_strided_*
  if (!A || !B || !C) { set_last_error("gemm_strided_batched: null operand pointer"); return AM_ERR_INVALID; }
  // products large enough for the tensor-core / DMMA kernels go through the single-product entry one by one (each
  // fills the chip on its own); small ones — the case batching exists for — share ONE launch (blockIdx.z = batch)
  const double flops = 2.0 * (double)M * (double)N * (double)K;
  const bool big_each = M >= 256 && N >= 256 && K >= 256 && flops >= 3.0e9;
  if (big_each && (std::is_same<T, float>::value || std::is_same<T, double>::value)) {
    for (int64_t b = 0; b < batch; b++) {
      int rc;
      if constexpr (std::is_same<T, float>::value)
        rc = am_gemm_strided_f32((am_stream_t)st, M, N, K, alpha, A + b * bsA, rsA, csA, B + b * bsB, rsB, csB, beta, C + b * bsC, rsC, csC);
      else if constexpr (std::is_same<T, double>::value)
        rc = am_gemm_strided_f64((am_stream_t)st, M, N, K, alpha, A + b * bsA, rsA, csA, B + b * bsB, rsB, csB, beta, C + b * bsC, rsC, csC);
      else rc = AM_ERR_INVALID;
      if (rc) return rc;
    }
    return AM_OK;
  }
  return gemm_simt_batched<T>(st, batch, M, N, K, alpha, A, rsA, csA, bsA, B, rsB, csB, bsB, beta, C, rsC, csC, bsC);
}

}  // namespace am

using namespace am;

extern "C" {

#define DEF_BATCHED(SUF, T)                                                                                       \
  int am_gemm_strided_batched_##SUF(am_stream_t s, int64_t batch, int64_t M, int64_t N, int64_t K, T alpha, const T* A, \
                                    int64_t rsA, int64_t csA, int64_t bsA, const T* B, int64_t rsB, int64_t csB,   \
                                    int64_t bsB, T beta, T* C, int64_t rsC, int64_t csC, int64_t bsC) {            \
    return gemm_batched<T>((cudaStream_t)s, batch, M, N, K, alpha, A, rsA, csA, bsA, B, rsB, csB, bsB, beta, C, rsC, csC, bsC); \
  }
DEF_BATCHED(f32, float)
DEF_BATCHED(f64, double)
DEF_BATCHED(i32, int32_t)
DEF_BATCHED(i64, int64_t)

#define DEF_CONV_STRIDED(SUF, T)                                                                                  \
  int am_conv2d_forward_strided_##SUF(am_stream_t s, const am_conv2d_desc* d, const T* in, const int64_t* is,     \
                                      const T* k, const int64_t* ks, const T* bias, int64_t bias_stride, T* out,  \
                                      const int64_t* os, int activation) {                                        \
    if (!d) { set_last_error("conv2d_forward_strided: null descriptor"); return AM_ERR_INVALID; }                \
    if (activation != AM_ACT_NONE && activation != AM_ACT_RELU) {                                                 \
      set_last_error("conv2d_forward_strided: unknown activation %d", activation); return AM_ERR_INVALID;         \
    }                                                                                                             \
    return conv_forward_strided<T>((cudaStream_t)s, *d, in, is, k, ks, bias, bias_stride, out, os, activation);   \
  }                                                                                                               \
  int am_conv2d_backward_strided_##SUF(am_stream_t s, const am_conv2d_desc* d, const T* in, const int64_t* is,    \
                                       const T* k, const int64_t* ks, const T* go, const int64_t* gos, T* gi,     \
                                       const int64_t* gis, T* gk, const int64_t* gks, T* gb, int64_t gb_stride) { \
    if (!d) { set_last_error("conv2d_backward_strided: null descriptor"); return AM_ERR_INVALID; }               \
    return conv_backward_strided<T>((cudaStream_t)s, *d, in, is, k, ks, go, gos, gi, gis, gk, gks, gb, gb_stride); \
  }
DEF_CONV_STRIDED(f32, float)
DEF_CONV_STRIDED(f64, double)
DEF_CONV_STRIDED(i32, int32_t)
DEF_CONV_STRIDED(i64, int64_t)

int am_memcpy2d_async(am_stream_t s, void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width,
                      int64_t rows, int kind) {
  if (width < 0 || rows < 0 || dst_pitch < width || src_pitch < width || kind < 1 || kind > 3) {
    set_last_error("am_memcpy2d_async: bad argument"); return AM_ERR_INVALID;
  }
  if (width == 0 || rows == 0) return AM_OK;
  if (!dst || !src) { set_last_error("am_memcpy2d_async: null pointer"); return AM_ERR_INVALID; }
  const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  AM_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width, (size_t)rows, k, (cudaStream_t)s));
  return AM_OK;
}

int64_t am_packed_floats_f32(int64_t R, int64_t K) { return (R > 0 && K > 0) ? packed_floats_f32(R, K) : 0; }
int am_pack_f32_a_into(am_stream_t s, int64_t M, int64_t K, const float* A, int64_t rsA, int64_t csA, float* planes,
                       am_packed_f32** out) {
  return pack_f32_view((cudaStream_t)s, M, K, A, rsA, csA, planes, (void**)out);
}
int am_pack_f32_b_into(am_stream_t s, int64_t K, int64_t N, const float* B, int64_t rsB, int64_t csB, float* planes,
                       am_packed_f32** out) {
  return pack_f32_view((cudaStream_t)s, N, K, B, csB, rsB, planes, (void**)out);
}
int am_packed_wrap_f32(int64_t R, int64_t K, float* planes, am_packed_f32** out) {
  return packed_wrap_f32(R, K, planes, (void**)out);
}

}  // extern "C"
