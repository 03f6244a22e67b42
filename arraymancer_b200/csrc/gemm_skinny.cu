// Skinny GEMMs — one dimension of the product at most 16 (a few rows of A against a large B, a large A against a few
// columns of B; N = 1 is the matrix-vector product the reference serves with gemv behind `*`,
// tensor/operators_blas_l2l3.nim:36-56).  These products are DRAM-bound: 2*S flop per element of the large operand
// (S <= 16) is below the FFMA ridge of the chip, so the kernel's job is to stream the large operand ONCE at HBM speed:
//   * 128-bit loads along the large operand's unit-stride dimension, 8 of them in flight per thread;
//   * the thin operand is staged in shared memory per k-chunk (it is re-read by every CTA, from L2);
//   * split-K over grid.y so that a [K x 16384] operand still yields ~4 CTAs per SM; partial sums are reduced in a
//     fixed order by a second (tiny) kernel that applies alpha / beta — deterministic, beta == 0 never reads C.
// Canonical form: OUT[s][g] = sum_k THIN[s][k] * BIG[k][g], s < S <= 16.  M <= 16 maps THIN = A, BIG = B; N <= 16 maps
// THIN = B^T, BIG = A^T (strides swapped, no copies).
//   skinny_kn_kernel : BIG has unit stride along g (row-major B, column-major A): a thread owns V consecutive g.
//   skinny_nk_kernel : BIG has unit stride along k (row-major A, transposed B): a warp owns 4 values of g, its lanes
//                      walk k with 128-bit loads, a fixed butterfly adds the 32 lane sums.
// Everything else (no unit stride, misaligned views) stays on the general SIMT kernel.
#include "am_common.cuh"
#include "gemm_dispatch.h"

namespace am {

template <class T>
struct SkinnyArgs {
  const T* thin; int64_t th_s, th_k;       // THIN[s][k]
  const T* big; int64_t bg_k, bg_g;        // BIG[k][g]
  T* out; int64_t o_s, o_g;                // OUT[s][g]
  T* part;                                 // [splits][S][Gd] partial sums (splits > 1), else null
  int64_t Gd, K, k_per_split;
  int S;                                   // valid thin rows (<= template S)
  T alpha, beta;
};

template <class T> struct VecOf { using type = int4; };
template <> struct VecOf<double> { using type = longlong2; };
template <> struct VecOf<int64_t> { using type = longlong2; };

// ------------------------------------------------------------------ BIG contiguous along g
template <class T, int S>
__global__ void __launch_bounds__(128) skinny_kn_kernel(const SkinnyArgs<T> a) {
  constexpr int V = 16 / (int)sizeof(T), KC = 128, U = 8;
  using Vec = typename VecOf<T>::type;
  __shared__ __align__(16) T ths[KC][S];
  const int tid = threadIdx.x;
  const int64_t g0 = ((int64_t)blockIdx.x * 128 + tid) * V;
  const bool g_ok = g0 < a.Gd;                      // Gd % V == 0 (checked by the host)
  const int64_t kb = (int64_t)blockIdx.y * a.k_per_split;
  const int64_t ke = (kb + a.k_per_split < a.K) ? kb + a.k_per_split : a.K;
  T acc[S][V];
#pragma unroll
  for (int s = 0; s < S; s++)
#pragma unroll
    for (int v = 0; v < V; v++) acc[s][v] = T(0);
  const T* bp = a.big + g0;
  // the U loads of the NEXT group of k steps are issued before the current group is multiplied — also across the barriers
  // of the thin-operand staging, so the DRAM stream never drains at a chunk boundary
  union VB { Vec q; T e[V]; };
  VB nxt[U];
  auto load_grp = [&](int64_t k0, VB (&dst)[U]) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (g_ok && k0 + u < ke) dst[u].q = __ldg(reinterpret_cast<const Vec*>(bp + (k0 + u) * a.bg_k));
      else {
#pragma unroll
        for (int v = 0; v < V; v++) dst[u].e[v] = T(0);
      }
    }
  };
  constexpr bool PF = S <= 4;                          // S = 8 / 16: the second buffer costs occupancy (measured: 0.62 -> 0.48)
  if (PF) load_grp(kb, nxt);
  for (int64_t kc = kb; kc < ke; kc += KC) {
    __syncthreads();
    for (int i = tid; i < KC * S; i += 128) {
      const int k = i / S, s = i - k * S;
      ths[k][s] = (s < a.S && kc + k < ke) ? a.thin[s * a.th_s + (kc + k) * a.th_k] : T(0);
    }
    __syncthreads();
    if (!PF && !g_ok) continue;
    const int kn = (int)((ke - kc < KC) ? ke - kc : KC);
    for (int k = 0; k < kn; k += U) {
      VB vb[U];
      if constexpr (PF) {
#pragma unroll
        for (int u = 0; u < U; u++) vb[u].q = nxt[u].q;
        load_grp(kc + k + U, nxt);                         // k steps past ke load nothing (zeros)
      } else {
#pragma unroll
        for (int u = 0; u < U; u++) {
          if (k + u < kn) vb[u].q = __ldg(reinterpret_cast<const Vec*>(bp + (kc + k + u) * a.bg_k));
          else {
#pragma unroll
            for (int v = 0; v < V; v++) vb[u].e[v] = T(0);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        T tv[S];
        if constexpr (S * sizeof(T) >= 16) {
#pragma unroll
          for (int q = 0; q < S / V; q++) {
            union { Vec w; T e[V]; } t; t.w = *reinterpret_cast<const Vec*>(&ths[k + u][q * V]);
#pragma unroll
            for (int v = 0; v < V; v++) tv[q * V + v] = t.e[v];
          }
        } else {
#pragma unroll
          for (int s = 0; s < S; s++) tv[s] = ths[k + u][s];
        }
#pragma unroll
        for (int s = 0; s < S; s++)
#pragma unroll
          for (int v = 0; v < V; v++) acc[s][v] = mac<T>(tv[s], vb[u].e[v], acc[s][v]);
      }
    }
  }
  if (!g_ok) return;
  if (a.part) {
    T* p = a.part + (int64_t)blockIdx.y * S * a.Gd + g0;
#pragma unroll
    for (int s = 0; s < S; s++) {
      union { Vec q; T e[V]; } o;
#pragma unroll
      for (int v = 0; v < V; v++) o.e[v] = acc[s][v];
      *reinterpret_cast<Vec*>(p + (int64_t)s * a.Gd) = o.q;
    }
  } else {
#pragma unroll
    for (int s = 0; s < S; s++)
      if (s < a.S)
#pragma unroll
        for (int v = 0; v < V; v++) {
          T* pc = a.out + s * a.o_s + (g0 + v) * a.o_g;
          *pc = epilogue_value<T>(a.alpha, acc[s][v], a.beta, a.beta != T(0) ? *pc : T(0));
        }
  }
}

// ------------------------------------------------------------------ BIG contiguous along k
template <class T> __device__ __forceinline__ T shfl_xor_t(T v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
template <> __device__ __forceinline__ int64_t shfl_xor_t<int64_t>(int64_t v, int o) {
  return (int64_t)__shfl_xor_sync(0xffffffffu, (long long)v, o);
}

template <class T, int S>
__global__ void __launch_bounds__(256) skinny_nk_kernel(const SkinnyArgs<T> a) {
  constexpr int V = 16 / (int)sizeof(T), R = 4, KC = 32 * V * 2;
  using Vec = typename VecOf<T>::type;
  __shared__ __align__(16) T ths[S][KC];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t g0 = ((int64_t)blockIdx.x * 8 + warp) * R;
  const int64_t kb = (int64_t)blockIdx.y * a.k_per_split;
  const int64_t ke = (kb + a.k_per_split < a.K) ? kb + a.k_per_split : a.K;
  T acc[R][S];
#pragma unroll
  for (int r = 0; r < R; r++)
#pragma unroll
    for (int s = 0; s < S; s++) acc[r][s] = T(0);
  const T* rowp[R];
  bool rok[R];
#pragma unroll
  for (int r = 0; r < R; r++) { rok[r] = g0 + r < a.Gd; rowp[r] = a.big + (rok[r] ? (g0 + r) : 0) * a.bg_g; }
  // the large operand of chunk c + 1 is requested before chunk c is multiplied: the loads' DRAM latency hides behind
  // S * R * V * 2 FMAs per lane instead of being paid after every barrier (first version: 0.27 of the HBM roof at N = 16)
  union VB { Vec q; T e[V]; };
  VB nxt[2][R];
  auto load_big = [&](int64_t kc, VB (&dst)[2][R]) {
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int64_t k = kc + (j * 32 + lane) * V;        // K % V == 0 and split slices are multiples of KC: k + V <= ke or k >= ke
#pragma unroll
      for (int r = 0; r < R; r++) {
        if (rok[r] && k < ke) dst[j][r].q = __ldg(reinterpret_cast<const Vec*>(rowp[r] + k));
        else {
#pragma unroll
          for (int v = 0; v < V; v++) dst[j][r].e[v] = T(0);
        }
      }
    }
  };
  constexpr bool PF = sizeof(T) == 4;                 // 8-byte types: the accumulators alone take 128 registers, no room to prefetch
  if (PF && kb < ke) load_big(kb, nxt);
  for (int64_t kc = kb; kc < ke; kc += KC) {
    __syncthreads();
    for (int i = tid; i < S * KC; i += 256) {
      const int s = i / KC, k = i - s * KC;
      ths[s][k] = (s < a.S && kc + k < ke) ? a.thin[s * a.th_s + (kc + k) * a.th_k] : T(0);
    }
    VB vb[2][R];
    if constexpr (PF) {
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int r = 0; r < R; r++) vb[j][r].q = nxt[j][r].q;
      if (kc + KC < ke) load_big(kc + KC, nxt);
    }
    __syncthreads();
    if constexpr (!PF) load_big(kc, vb);
#pragma unroll
    for (int j = 0; j < 2; j++) {
#pragma unroll
      for (int s = 0; s < S; s++) {
        union { Vec w; T e[V]; } t; t.w = *reinterpret_cast<const Vec*>(&ths[s][(j * 32 + lane) * V]);
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
          for (int v = 0; v < V; v++) acc[r][s] = mac<T>(vb[j][r].e[v], t.e[v], acc[r][s]);
      }
    }
  }
  // fixed butterfly over the 32 lanes
#pragma unroll
  for (int r = 0; r < R; r++)
#pragma unroll
    for (int s = 0; s < S; s++) {
      T v = acc[r][s];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = add_nocontract<T>(v, shfl_xor_t<T>(v, o));
      acc[r][s] = v;
    }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < R; r++) {
      if (!rok[r]) continue;
#pragma unroll
      for (int s = 0; s < S; s++) {
        if (a.part) a.part[((int64_t)blockIdx.y * S + s) * a.Gd + g0 + r] = acc[r][s];
        else if (s < a.S) {
          T* pc = a.out + s * a.o_s + (g0 + r) * a.o_g;
          *pc = epilogue_value<T>(a.alpha, acc[r][s], a.beta, a.beta != T(0) ? *pc : T(0));
        }
      }
    }
  }
}

// partial sums of the split-K slices, added in ascending slice order, then the alpha / beta epilogue
template <class T>
__global__ void __launch_bounds__(256) skinny_reduce_kernel(const T* __restrict__ part, int splits, int SP, int S, int64_t Gd,
                                                            T* out, int64_t o_s, int64_t o_g, T alpha, T beta) {
  const int64_t total = (int64_t)S * Gd;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / Gd, g = i - s * Gd;
    T acc = T(0);
    for (int z = 0; z < splits; z++) acc = add_nocontract<T>(acc, part[((int64_t)z * SP + s) * Gd + g]);
    T* pc = out + s * o_s + g * o_g;
    *pc = epilogue_value<T>(alpha, acc, beta, beta != T(0) ? *pc : T(0));
  }
}

template <class T, int S>
static int launch_skinny(cudaStream_t st, SkinnyArgs<T> a, bool kn, int splits) {
  constexpr int V = 16 / (int)sizeof(T);
  if (kn) {
    dim3 grid((unsigned)ceil_div(a.Gd, 128 * V), (unsigned)splits);
    skinny_kn_kernel<T, S><<<grid, 128, 0, st>>>(a);
  } else {
    dim3 grid((unsigned)ceil_div(a.Gd, 32), (unsigned)splits);
    skinny_nk_kernel<T, S><<<grid, 256, 0, st>>>(a);
  }
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  return AM_OK;
}

// resident CTAs of the chosen kernel on the whole chip (sizes the split-K grid, see gemm_skinny)
template <class T, int S>
static int64_t skinny_slots(bool kn) {
  int per_sm = 0;
  cudaError_t e = kn ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, skinny_kn_kernel<T, S>, 128, 0)
                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, skinny_nk_kernel<T, S>, 256, 0);
  if (e != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
  return (int64_t)per_sm * sm_count();
}

// *done = false: the shape / layout is not a skinny case this file handles (caller continues with the general kernel).
template <class T>
int gemm_skinny(cudaStream_t st, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t rsA, int64_t csA, const T* B,
                int64_t rsB, int64_t csB, T beta, T* C, int64_t rsC, int64_t csC, bool* done) {
  *done = false;
  constexpr int V = 16 / (int)sizeof(T);
  const bool m_small = M <= 16 && M <= N;
  const bool n_small = !m_small && N <= 16;
  if (!m_small && !n_small) return AM_OK;
  SkinnyArgs<T> a{};
  if (m_small) { a.thin = A; a.th_s = rsA; a.th_k = csA; a.big = B; a.bg_k = rsB; a.bg_g = csB; a.o_s = rsC; a.o_g = csC; a.S = (int)M; a.Gd = N; }
  else { a.thin = B; a.th_s = csB; a.th_k = rsB; a.big = A; a.bg_k = csA; a.bg_g = rsA; a.o_s = csC; a.o_g = rsC; a.S = (int)N; a.Gd = M; }
  a.out = C; a.K = K; a.alpha = alpha; a.beta = beta;
  // worth it only when the large operand is large: K * Gd elements streamed once
  if ((double)K * (double)a.Gd < 1048576.0 || a.Gd < 256) return AM_OK;
  const bool aligned = (reinterpret_cast<uintptr_t>(a.big) & 15) == 0;
  bool kn;
  if (a.bg_g == 1 && aligned && a.bg_k % V == 0 && a.Gd % V == 0) kn = true;
  else if (a.bg_k == 1 && aligned && a.bg_g % V == 0 && K % V == 0) kn = false;
  else return AM_OK;
  // split K so that the grid holds ~4 CTAs per SM; slices are multiples of the k-chunk of the kernel
  const int64_t KC = kn ? 128 : 32 * V * 2;
  const int64_t ctas_g = kn ? ceil_div(a.Gd, 128 * V) : ceil_div(a.Gd, 32);
  const int SP = a.S <= 1 ? 1 : a.S <= 2 ? 2 : a.S <= 4 ? 4 : a.S <= 8 ? 8 : 16;
  int64_t slots;
  switch (SP) {
    case 1: slots = skinny_slots<T, 1>(kn); break;
    case 2: slots = skinny_slots<T, 2>(kn); break;
    case 4: slots = skinny_slots<T, 4>(kn); break;
    case 8: slots = skinny_slots<T, 8>(kn); break;
    default: slots = skinny_slots<T, 16>(kn); break;
  }
  // split-K depth, measured at 16384^2 (profiles/r02 bench, GB/s of the large operand): what matters is the number of loads in
  // flight, so the thin kernels (S = 1: few registers, up to 16 CTAs per SM) fill every resident slot — M = 1: 3.0 -> 4.6 TB/s,
  // N = 1: 5.9 -> 6.2 TB/s with one full wave — while the register-heavy S > 1 kernels keep the ~4 CTAs per SM of round 1
  // (one exact wave of fewer CTAs measured 12 % slower there).
  const int64_t per_sm = slots / sm_count();
  int64_t splits;
  if (SP == 1 && !kn) splits = slots / ctas_g;
  else splits = ceil_div((SP == 1 && per_sm > 4 ? per_sm : 4) * (int64_t)sm_count(), ctas_g);
  const int64_t max_splits = ceil_div(K, 4 * KC);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  a.k_per_split = round_up(ceil_div(K, splits), KC);
  splits = ceil_div(K, a.k_per_split);                 // (rounding the slice up can only lower the count)
  a.part = nullptr;
  if (splits > 1) {
    void* p = nullptr;
    int rc = workspace(kWsSkinny, (size_t)(splits * SP * a.Gd) * sizeof(T), &p);
    if (rc) return rc;
    a.part = (T*)p;
  }
  int rc;
  switch (SP) {
    case 1: rc = launch_skinny<T, 1>(st, a, kn, (int)splits); break;
    case 2: rc = launch_skinny<T, 2>(st, a, kn, (int)splits); break;
    case 4: rc = launch_skinny<T, 4>(st, a, kn, (int)splits); break;
    case 8: rc = launch_skinny<T, 8>(st, a, kn, (int)splits); break;
    default: rc = launch_skinny<T, 16>(st, a, kn, (int)splits); break;
  }
  if (rc) return rc;
  if (splits > 1) {
    const int64_t total = (int64_t)a.S * a.Gd;
    int64_t blocks = ceil_div(total, 256);
    if (blocks > 8 * (int64_t)sm_count()) blocks = 8 * (int64_t)sm_count();
    skinny_reduce_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(a.part, (int)splits, SP, a.S, a.Gd, a.out, a.o_s, a.o_g, alpha, beta);
    g_launch_count++;
    AM_CUDA_TRY(cudaGetLastError());
  }
  *done = true;
  return AM_OK;
}

#define INST_SKINNY(T)                                                                                                     \
  template int gemm_skinny<T>(cudaStream_t, int64_t, int64_t, int64_t, T, const T*, int64_t, int64_t, const T*, int64_t,   \
                              int64_t, T, T*, int64_t, int64_t, bool*);
INST_SKINNY(float)
INST_SKINNY(double)
INST_SKINNY(int32_t)
INST_SKINNY(int64_t)

}  // namespace am
