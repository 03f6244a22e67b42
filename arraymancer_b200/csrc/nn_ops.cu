// The small HBM-bound operators that sit between the contractions of the reference's LeNet example (SURVEY §8f rows
// 1-3), so that a forward + backward step can stay resident on the device: ReLU, MaxPool2D, the bias / bias-gradient
// passes of the linear layer and sparse softmax cross-entropy.  Semantics follow the reference line by line:
//   relu / relu_backward                  nn_primitives/nnp_activation.nim:35-36, 65-70
//   maxpool2d / maxpool2d_backward        nn_primitives/nnp_maxpooling.nim:19-83
//   linear (+ bias), linear_backward      nn_primitives/nnp_linear.nim:20-66   (the products go through gemm_strided)
//   sparse_softmax_cross_entropy (+ bwd)  nn_primitives/nnp_softmax_cross_entropy.nim:100-178, 219-252,
//                                         private/p_logsumexp.nim:13-23 (streaming max / sum-exp)
// All kernels are deterministic (fixed reduction orders, no floating-point atomics).
#include <cfloat>
#include <climits>

#include "am_common.cuh"
#include "gemm_dispatch.h"

namespace am {

// the products of the linear layer go through the public strided-GEMM entries (same dispatch as CudaTensor `*`)
template <class T>
static int gemm_strided(cudaStream_t st, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t rsA, int64_t csA,
                        const T* B, int64_t rsB, int64_t csB, T beta, T* C, int64_t rsC, int64_t csC);
template <>
int gemm_strided<float>(cudaStream_t st, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t rsA,
                        int64_t csA, const float* B, int64_t rsB, int64_t csB, float beta, float* C, int64_t rsC, int64_t csC) {
  return am_gemm_strided_f32((am_stream_t)st, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC);
}
template <>
int gemm_strided<double>(cudaStream_t st, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t rsA,
                         int64_t csA, const double* B, int64_t rsB, int64_t csB, double beta, double* C, int64_t rsC, int64_t csC) {
  return am_gemm_strided_f64((am_stream_t)st, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC);
}

// ------------------------------------------------------------------ ReLU
// max(0, x) with Nim's max: "if x <= 0: 0 else: x" — a NaN input stays NaN.
template <class T>
__global__ void relu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const T v = x[i];
    y[i] = (v <= T(0)) ? T(0) : v;
  }
}
// relu_backward(gradient, cached): cached <= 0 ? 0 : gradient
template <class T>
__global__ void relu_bwd_kernel(const T* __restrict__ grad, const T* __restrict__ cached, T* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (cached[i] <= T(0)) ? T(0) : grad[i];
}
// 128-bit variants for the aligned float32 case (the LeNet activations: 47-189 MB per pass)
__global__ void relu_fwd_f32x4_kernel(const float4* __restrict__ x, float4* __restrict__ y, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = x[i];
    v.x = (v.x <= 0.f) ? 0.f : v.x; v.y = (v.y <= 0.f) ? 0.f : v.y;
    v.z = (v.z <= 0.f) ? 0.f : v.z; v.w = (v.w <= 0.f) ? 0.f : v.w;
    y[i] = v;
  }
}
__global__ void relu_bwd_f32x4_kernel(const float4* __restrict__ g, const float4* __restrict__ c, float4* __restrict__ o, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 gv = g[i], cv = c[i];
    o[i] = make_float4((cv.x <= 0.f) ? 0.f : gv.x, (cv.y <= 0.f) ? 0.f : gv.y, (cv.z <= 0.f) ? 0.f : gv.z, (cv.w <= 0.f) ? 0.f : gv.w);
  }
}

static unsigned grid_for(int64_t n, int threads) {
  int64_t b = ceil_div(n, (int64_t)threads);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}

template <class T>
int relu_forward(cudaStream_t st, int64_t n, const T* x, T* y) {
  if (n < 0 || (n > 0 && (!x || !y))) { set_last_error("relu_forward: bad argument"); return AM_ERR_INVALID; }
  if (n == 0) return AM_OK;
  if constexpr (std::is_same<T, float>::value) {
    if (n % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
      relu_fwd_f32x4_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>((const float4*)x, (float4*)y, n / 4);
      g_launch_count++;
      AM_CUDA_TRY(cudaGetLastError());
      return AM_OK;
    }
  }
  relu_fwd_kernel<T><<<grid_for(n, 256), 256, 0, st>>>(x, y, n);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  return AM_OK;
}
template <class T>
int relu_backward(cudaStream_t st, int64_t n, const T* grad, const T* cached, T* out) {
  if (n < 0 || (n > 0 && (!grad || !cached || !out))) { set_last_error("relu_backward: bad argument"); return AM_ERR_INVALID; }
  if (n == 0) return AM_OK;
  if constexpr (std::is_same<T, float>::value) {
    if (n % 4 == 0 && ((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(cached) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
      relu_bwd_f32x4_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>((const float4*)grad, (const float4*)cached, (float4*)out, n / 4);
      g_launch_count++;
      AM_CUDA_TRY(cudaGetLastError());
      return AM_OK;
    }
  }
  relu_bwd_kernel<T><<<grid_for(n, 256), 256, 0, st>>>(grad, cached, out, n);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  return AM_OK;
}

// ------------------------------------------------------------------ MaxPool2D
template <class T> __device__ __forceinline__ T lowest_value();
template <> __device__ __forceinline__ float lowest_value<float>() { return -INFINITY; }       // Nim: low(float32) = -Inf
template <> __device__ __forceinline__ double lowest_value<double>() { return -INFINITY; }
template <> __device__ __forceinline__ int32_t lowest_value<int32_t>() { return INT32_MIN; }
template <> __device__ __forceinline__ int64_t lowest_value<int64_t>() { return INT64_MIN; }

// one thread per output element (w fastest: coalesced stores, overlapping reads served by L1/L2); window scanned in
// (ph, pw) order with a strict '>' so the first maximum wins; no valid element -> (low(T), low(int))
template <class T>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t* __restrict__ idx, int64_t total,
                                   int C, int H, int W, int Ho, int Wo, int kH, int kW, int padH, int padW, int sH, int sW) {
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(o % Wo);
    const int64_t t1 = o / Wo;
    const int h = (int)(t1 % Ho);
    const int64_t nc = t1 / Ho;                             // n * C + c
    const T* plane = x + nc * (int64_t)H * W;
    T best = lowest_value<T>();
    int64_t arg = INT64_MIN;
    for (int ph = 0; ph < kH; ph++) {
      const int row = h * sH + ph - padH;
      if (row < 0 || row >= H) continue;
      for (int pw = 0; pw < kW; pw++) {
        const int col = w * sW + pw - padW;
        if (col < 0 || col >= W) continue;
        const T v = plane[row * W + col];
        if (v > best) { best = v; arg = nc * (int64_t)H * W + (int64_t)row * W + col; }
      }
    }
    y[o] = best;
    idx[o] = arg;
  }
}
// 2x2 / stride 2 / no padding, float32, W % 4 == 0 (LeNet's pools): a thread takes two neighbouring windows from two
// 128-bit loads and writes its two results with one 64-bit and one 128-bit store; same scan order and tie rule
__global__ void maxpool_2x2_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t* __restrict__ idx,
                                       int64_t npairs, int H, int W, int Ho, int Wo) {
  const int wp = Wo >> 1;                                   // output pairs per row
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < npairs; t += (int64_t)gridDim.x * blockDim.x) {
    const int pw = (int)(t % wp);
    const int64_t t1 = t / wp;
    const int h = (int)(t1 % Ho);
    const int64_t nc = t1 / Ho;
    const int64_t base = nc * (int64_t)H * W + (int64_t)(2 * h) * W + 4 * pw;
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(x + base));
    const float4 r1 = __ldg(reinterpret_cast<const float4*>(x + base + W));
    float b0 = -INFINITY, b1 = -INFINITY;
    int64_t a0 = INT64_MIN, a1 = INT64_MIN;
    if (r0.x > b0) { b0 = r0.x; a0 = base; }
    if (r0.y > b0) { b0 = r0.y; a0 = base + 1; }
    if (r1.x > b0) { b0 = r1.x; a0 = base + W; }
    if (r1.y > b0) { b0 = r1.y; a0 = base + W + 1; }
    if (r0.z > b1) { b1 = r0.z; a1 = base + 2; }
    if (r0.w > b1) { b1 = r0.w; a1 = base + 3; }
    if (r1.z > b1) { b1 = r1.z; a1 = base + W + 2; }
    if (r1.w > b1) { b1 = r1.w; a1 = base + W + 3; }
    const int64_t o = (nc * Ho + h) * (int64_t)Wo + 2 * pw;
    *reinterpret_cast<float2*>(y + o) = make_float2(b0, b1);
    *reinterpret_cast<longlong2*>(idx + o) = make_longlong2(a0, a1);
  }
}

// gradInput[max_indices[i]] = gradOutput[i] (assignment, nnp_maxpooling.nim:82-83).  With overlapping windows the
// serial reference keeps the LAST i: pass 1 records the largest i per input slot, pass 2 lets only that i write.
__global__ void maxpool_owner_kernel(const int64_t* __restrict__ idx, long long* __restrict__ owner, int64_t n_out, int64_t n_in) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_out; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx[i];
    if (j >= 0 && j < n_in) atomicMax(&owner[j], (long long)i);
  }
}
// relu_cached != null: the relu_backward of a conv -> relu -> maxpool block (nnp_activation.nim:65-70, `cached <= 0 ? 0 :
// gradient`) is applied to the one element per window that receives a gradient — one gathered load per POOLED output
// instead of an element-wise pass (read gradient, read cached, write) over the whole activation tensor.
template <class T>
__global__ void maxpool_bwd_kernel(const int64_t* __restrict__ idx, const long long* __restrict__ owner, const T* __restrict__ go,
                                   const T* __restrict__ relu_cached, T* __restrict__ gi, int64_t n_out, int64_t n_in) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_out; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx[i];
    if (j >= 0 && j < n_in && (owner == nullptr || owner[j] == (long long)i)) {
      T g = go[i];
      if (relu_cached != nullptr && relu_cached[j] <= T(0)) g = T(0);
      gi[j] = g;
    }
  }
}

template <class T>
int maxpool2d_forward(cudaStream_t st, int64_t N, int64_t C, int64_t H, int64_t W, int64_t kH, int64_t kW, int64_t padH,
                      int64_t padW, int64_t sH, int64_t sW, const T* x, T* y, int64_t* idx) {
  if (N < 0 || C < 1 || H < 1 || W < 1 || kH < 1 || kW < 1 || sH < 1 || sW < 1 || padH < 0 || padW < 0) {
    set_last_error("maxpool2d_forward: invalid geometry"); return AM_ERR_INVALID;
  }
  const int64_t Ho = (H + 2 * padH - kH) / sH + 1, Wo = (W + 2 * padW - kW) / sW + 1;     // nnp_maxpooling.nim:37-38
  if (Ho < 1 || Wo < 1 || H >= (1 << 30) || W >= (1 << 30) || C >= (1ll << 31)) { set_last_error("maxpool2d_forward: invalid geometry"); return AM_ERR_INVALID; }
  const int64_t total = N * C * Ho * Wo;
  if (total == 0) return AM_OK;
  if (!x || !y || !idx) { set_last_error("maxpool2d_forward: null pointer"); return AM_ERR_INVALID; }
  if constexpr (std::is_same<T, float>::value) {
    if (kH == 2 && kW == 2 && sH == 2 && sW == 2 && padH == 0 && padW == 0 && W % 4 == 0 && H % 2 == 0 &&
        (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0 && (reinterpret_cast<uintptr_t>(idx) & 15) == 0) {
      maxpool_2x2_f32_kernel<<<grid_for(total / 2, 256), 256, 0, st>>>(x, y, idx, total / 2, (int)H, (int)W, (int)Ho, (int)Wo);
      g_launch_count++;
      AM_CUDA_TRY(cudaGetLastError());
      return AM_OK;
    }
  }
  maxpool_fwd_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(x, y, idx, total, (int)C, (int)H, (int)W, (int)Ho, (int)Wo,
                                                             (int)kH, (int)kW, (int)padH, (int)padW, (int)sH, (int)sW);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  return AM_OK;
}

template <class T>
int maxpool2d_backward(cudaStream_t st, int64_t n_in, int64_t n_out, const int64_t* idx, const T* go, const T* relu_cached, T* gi,
                       int windows_overlap) {
  if (n_in < 0 || n_out < 0 || (n_in > 0 && !gi) || (n_out > 0 && (!idx || !go))) { set_last_error("maxpool2d_backward: bad argument"); return AM_ERR_INVALID; }
  if (n_in == 0) return AM_OK;
  AM_CUDA_TRY(cudaMemsetAsync(gi, 0, (size_t)n_in * sizeof(T), st));                      // zeros(cached_input_shape)
  if (n_out == 0) return AM_OK;
  long long* owner = nullptr;
  if (windows_overlap) {
    void* ws = nullptr;
    int rc = workspace(kWsNn, (size_t)n_in * sizeof(long long), &ws);
    if (rc) return rc;
    owner = (long long*)ws;
    AM_CUDA_TRY(cudaMemsetAsync(owner, 0x80, (size_t)n_in * sizeof(long long), st));      // very negative
    maxpool_owner_kernel<<<grid_for(n_out, 256), 256, 0, st>>>(idx, owner, n_out, n_in);
    g_launch_count++;
  }
  maxpool_bwd_kernel<T><<<grid_for(n_out, 256), 256, 0, st>>>(idx, owner, go, relu_cached, gi, n_out, n_in);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  return AM_OK;
}

// ------------------------------------------------------------------ linear layer: bias passes
// output +.= bias (row-major [batch, out]; bias [out]); nnp_linear.nim:28-29
template <class T>
__global__ void add_bias_rows_kernel(T* __restrict__ y, const T* __restrict__ bias, int64_t batch, int64_t out) {
  const int64_t total = batch * out;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = add_nocontract<T>(y[i], bias[i % out]);
}
// gradBias = sum(gradOutput, axis = 0) (nnp_linear.nim:52): block = 32 columns x 8 row lanes; each lane sums its rows
// in ascending order, the 8 partials are added in fixed order
template <class T>
__global__ void colsum_kernel(const T* __restrict__ g, T* __restrict__ out, int64_t batch, int64_t cols) {
  __shared__ T part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + tx;
  T acc = T(0);
  if (c < cols)
    for (int64_t r = ty; r < batch; r += 8) acc = add_nocontract<T>(acc, g[r * cols + c]);
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < cols) {
    T s = part[0][tx];
    for (int k = 1; k < 8; k++) s = add_nocontract<T>(s, part[k][tx]);
    out[c] = s;
  }
}

// y[batch, out] = x[batch, in] * W[out, in]^T (+ bias[out])
template <class T>
int linear_forward(cudaStream_t st, int64_t batch, int64_t in, int64_t out, const T* x, const T* w, const T* bias, T* y) {
  if (batch < 0 || in < 0 || out < 0) { set_last_error("linear_forward: negative extent"); return AM_ERR_INVALID; }
  if (batch == 0 || out == 0) return AM_OK;
  if (in > 0) {
    // y = x * W^T (W^T as a view: rs = 1, cs = in) with `+ bias` fused into the GEMM epilogue (the
    // reference adds it in a second pass, nnp_linear.nim:28-29): same value, rounded like the separate pass
    if constexpr (std::is_same<T, float>::value)
      return gemm_dispatch_f32(st, batch, out, in, 1.f, x, in, 1, w, 1, in, 0.f, y, out, 1, bias);
    else
      return gemm_dispatch_f64(st, batch, out, in, 1.0, x, in, 1, w, 1, in, 0.0, y, out, 1, bias);
  }
  AM_CUDA_TRY(cudaMemsetAsync(y, 0, (size_t)(batch * out) * sizeof(T), st));       // empty product: y = bias
  if (bias) {
    add_bias_rows_kernel<T><<<grid_for(batch * out, 256), 256, 0, st>>>(y, bias, batch, out);
    g_launch_count++;
    AM_CUDA_TRY(cudaGetLastError());
  }
  return AM_OK;
}
// gradInput = gO * W ; gradWeight = gO^T * x ; gradBias = sum(gO, axis 0); any output may be null
template <class T>
int linear_backward(cudaStream_t st, int64_t batch, int64_t in, int64_t out, const T* x, const T* w, const T* go, T* gi,
                    T* gw, T* gb) {
  if (batch < 0 || in < 0 || out < 0 || !go) { set_last_error("linear_backward: bad argument"); return AM_ERR_INVALID; }
  int rc;
  if (gi && batch > 0 && in > 0) {
    if (!w) { set_last_error("linear_backward: weight needed for gradInput"); return AM_ERR_INVALID; }
    if (out == 0) AM_CUDA_TRY(cudaMemsetAsync(gi, 0, (size_t)(batch * in) * sizeof(T), st));
    else if ((rc = gemm_strided<T>(st, batch, in, out, T(1), go, out, 1, w, in, 1, T(0), gi, in, 1))) return rc;
  }
  if (gw && out > 0 && in > 0) {
    if (!x) { set_last_error("linear_backward: input needed for gradWeight"); return AM_ERR_INVALID; }
    if (batch == 0) AM_CUDA_TRY(cudaMemsetAsync(gw, 0, (size_t)(out * in) * sizeof(T), st));
    else if ((rc = gemm_strided<T>(st, out, in, batch, T(1), go, 1, out, x, in, 1, T(0), gw, in, 1))) return rc;   // gO^T view
  }
  if (gb && out > 0) {
    colsum_kernel<T><<<(unsigned)ceil_div(out, (int64_t)32), 256, 0, st>>>(go, gb, batch, out);
    g_launch_count++;
    AM_CUDA_TRY(cudaGetLastError());
  }
  return AM_OK;
}

// ------------------------------------------------------------------ sparse softmax cross-entropy
template <class T> __device__ __forceinline__ T dev_exp(T v);
template <> __device__ __forceinline__ float dev_exp<float>(float v) { return expf(v); }
template <> __device__ __forceinline__ double dev_exp<double>(double v) { return exp(v); }
template <class T> __device__ __forceinline__ T dev_log(T v);
template <> __device__ __forceinline__ float dev_log<float>(float v) { return logf(v); }
template <> __device__ __forceinline__ double dev_log<double>(double v) { return log(v); }

// streaming max / sum-exp of one row, exactly the reference's single pass (p_logsumexp.nim:13-23)
template <class T>
__device__ __forceinline__ void row_max_sumexp(const T* row, int64_t cs, int64_t features, T* mx, T* se) {
  T m = -INFINITY, s = T(0);
  for (int64_t j = 0; j < features; j++) {
    const T v = row[j * cs];
    if (v <= m) s += dev_exp<T>(v - m);
    else { s = s * dev_exp<T>(m - v) + T(1); m = v; }
  }
  *mx = m; *se = s;
}
// one thread per sample (the class count of the reference's use is 10-1000): loss_i = ln(sumexp) + max - x[i, label]
template <class T>
__global__ void ssce_rows_kernel(const T* __restrict__ x, int64_t rs, int64_t cs, const int64_t* __restrict__ labels,
                                 int64_t batch, int64_t features, T* __restrict__ row_loss) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < batch; i += (int64_t)gridDim.x * blockDim.x) {
    const T* row = x + i * rs;
    T m, s;
    row_max_sumexp<T>(row, cs, features, &m, &s);
    const int64_t lab = labels[i];
    const T at = (lab >= 0 && lab < features) ? row[lab * cs] : T(NAN);
    row_loss[i] = dev_log<T>(s) + m - at;
  }
}
// mean of the row losses: one block, fixed-order tree (the reference adds them with an OpenMP atomic in no fixed order)
template <class T>
__global__ void mean_kernel(const T* __restrict__ v, int64_t n, T* __restrict__ out) {
  __shared__ T sh[256];
  T acc = T(0);
  for (int64_t i = threadIdx.x; i < n; i += 256) acc += v[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if ((int)threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0] / T(n);
}
// result[i, j] = grad * (softmax(x)[i, j] - [j == label_i]) / batch   (nnp_softmax_cross_entropy.nim:240-252)
template <class T>
__global__ void ssce_bwd_kernel(const T* __restrict__ x, int64_t rs, int64_t cs, const int64_t* __restrict__ labels, int64_t batch,
                                int64_t features, T grad, T* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < batch; i += (int64_t)gridDim.x * blockDim.x) {
    const T* row = x + i * rs;
    T m, s;
    row_max_sumexp<T>(row, cs, features, &m, &s);
    const int64_t lab = labels[i];
    for (int64_t j = 0; j < features; j++) {
      const T sm = dev_exp<T>(row[j * cs] - m) / s;
      out[i * features + j] = grad * (sm + (j == lab ? T(-1) : T(0))) / T(batch);
    }
  }
}

template <class T>
int ssce_forward(cudaStream_t st, int64_t batch, int64_t features, const T* x, int64_t rs, int64_t cs, const int64_t* labels,
                 T* loss_dev) {
  if (batch < 0 || features < 0 || !loss_dev) { set_last_error("sparse_softmax_cross_entropy: bad argument"); return AM_ERR_INVALID; }
  if (batch == 0) { AM_CUDA_TRY(cudaMemsetAsync(loss_dev, 0, sizeof(T), st)); return AM_OK; }        // returns 0 (:128-130)
  if (!x || !labels) { set_last_error("sparse_softmax_cross_entropy: null pointer"); return AM_ERR_INVALID; }
  void* ws = nullptr;
  int rc = workspace(kWsNn, (size_t)batch * sizeof(T), &ws);
  if (rc) return rc;
  ssce_rows_kernel<T><<<grid_for(batch, 128), 128, 0, st>>>(x, rs, cs, labels, batch, features, (T*)ws);
  mean_kernel<T><<<1, 256, 0, st>>>((const T*)ws, batch, loss_dev);
  g_launch_count += 2;
  AM_CUDA_TRY(cudaGetLastError());
  return AM_OK;
}
template <class T>
int ssce_backward(cudaStream_t st, int64_t batch, int64_t features, T grad, const T* x, int64_t rs, int64_t cs,
                  const int64_t* labels, T* out) {
  if (batch < 0 || features < 0) { set_last_error("sparse_softmax_cross_entropy_backward: bad argument"); return AM_ERR_INVALID; }
  if (batch == 0 || features == 0) return AM_OK;
  if (!x || !labels || !out) { set_last_error("sparse_softmax_cross_entropy_backward: null pointer"); return AM_ERR_INVALID; }
  ssce_bwd_kernel<T><<<grid_for(batch, 128), 128, 0, st>>>(x, rs, cs, labels, batch, features, grad, out);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  return AM_OK;
}

#define INST(T)                                                                                                          \
  template int relu_forward<T>(cudaStream_t, int64_t, const T*, T*);                                                     \
  template int relu_backward<T>(cudaStream_t, int64_t, const T*, const T*, T*);                                          \
  template int maxpool2d_forward<T>(cudaStream_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, \
                                    int64_t, int64_t, const T*, T*, int64_t*);                                           \
  template int maxpool2d_backward<T>(cudaStream_t, int64_t, int64_t, const int64_t*, const T*, const T*, T*, int);       \
  template int linear_forward<T>(cudaStream_t, int64_t, int64_t, int64_t, const T*, const T*, const T*, T*);             \
  template int linear_backward<T>(cudaStream_t, int64_t, int64_t, int64_t, const T*, const T*, const T*, T*, T*, T*);
INST(float)
INST(double)
#undef INST
template int ssce_forward<float>(cudaStream_t, int64_t, int64_t, const float*, int64_t, int64_t, const int64_t*, float*);
template int ssce_forward<double>(cudaStream_t, int64_t, int64_t, const double*, int64_t, int64_t, const int64_t*, double*);
template int ssce_backward<float>(cudaStream_t, int64_t, int64_t, float, const float*, int64_t, int64_t, const int64_t*, float*);
template int ssce_backward<double>(cudaStream_t, int64_t, int64_t, double, const double*, int64_t, int64_t, const int64_t*, double*);

}  // namespace am
