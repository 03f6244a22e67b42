// Strided SIMT GEMM for all four dtypes (int32/int64: IMAD, exact mod 2^n; f64: DFMA;
// f32: exact FFMA path used for small / skinny shapes and as the AM_F32_SIMT selector).
// Device-side replacement of laser gemm_strided (gemm.nim:192-273) for arbitrary
// (rowStride, colStride) views — negative, zero and transposed strides included.
#pragma once
#include "contract_simt.cuh"
#include "gemm_dispatch.h"

namespace am {


template <class T>
struct GemmCfgs {
  static constexpr int BK = 64 / (int)sizeof(T);
  using Big = SimtCfg<T, 128, 128, BK, 8, 8>;     // 256 threads, 8x8 register tile
  using Small = SimtCfg<T, 64, 64, BK, 4, 4>;     // 256 threads, 4x4 register tile
};
// int64: a multiply-accumulate is three dependent IMADs, shared memory is nowhere near the limit, so a
// 4x8 register tile (half the accumulator registers) with 512 threads doubles the resident warps per
// scheduler (latency hiding) at the same 128x128 CTA tile.
template <>
struct GemmCfgs<int64_t> {
  static constexpr int BK = 8;
  using Big = SimtCfg<int64_t, 128, 128, BK, 4, 8>;   // 512 threads, 4x8 register tile
  using Small = SimtCfg<int64_t, 64, 64, BK, 4, 4>;
};

// Pre-pass of the int64 GEMM (gemm_simt_i64.cu): is every element of A and of B representable in int32?
__global__ void i64_range_kernel(const int64_t* __restrict__ A, int64_t a_mn, int64_t a_k, int64_t M,
                                 const int64_t* __restrict__ B, int64_t b_mn, int64_t b_k, int64_t N, int64_t K,
                                 int* __restrict__ wide_flag);

template <>
struct GemmCfgs<int32_t> {
  static constexpr int BK = 16;
  using Big = SimtCfg<int32_t, 128, 128, BK, 4, 8>;   // 512 threads: IMAD issues at half the FFMA rate, occupancy wins
  using Small = SimtCfg<int32_t, 64, 64, BK, 4, 4>;
};

template <class T, class Cfg, bool Batched, int MA, int MB>
static int launch_mode(cudaStream_t st, dim3 grid, const StridedLoader<T>& la, const StridedLoader<T>& lb,
                       const StridedEpilogue<T>& epi, int64_t K, int a_kfast, int b_kfast, const int* wide_flag,
                       int64_t bsA, int64_t bsB, int64_t bsC) {
  contract_simt_kernel<T, Cfg, StridedLoader<T>, StridedLoader<T>, StridedEpilogue<T>, Batched, MA, MB>
      <<<grid, Cfg::NT, 0, st>>>(la, lb, epi, K, K, a_kfast, b_kfast, wide_flag, bsA, bsB, bsC);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  return AM_OK;
}

// staging mode of one operand (contract_simt.cuh): 128-bit accesses along its unit-stride dimension when the layout allows
template <class T>
static int stage_mode(const T* p, int64_t mn_stride, int64_t k_stride, int64_t batch_stride) {
  constexpr int V = 16 / (int)sizeof(T);
  if (!tuning(kTuneSimtVecLoad) || (reinterpret_cast<uintptr_t>(p) & 15) != 0 || batch_stride % V != 0) return kStageScalar;
  if (mn_stride == 1 && k_stride % V == 0) return kStageVecMN;
  if (k_stride == 1 && mn_stride % V == 0) return kStageVecK;
  return kStageScalar;
}

template <class T, class Cfg, bool Batched>
static int launch_cfg(cudaStream_t st, int64_t batch, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t a_mn,
                      int64_t a_k, int64_t bsA, const T* B, int64_t b_mn, int64_t b_k, int64_t bsB, T beta, T* C,
                      int64_t rsC, int64_t csC, int64_t bsC, const T* bias_m = nullptr, const T* bias_n = nullptr) {
  using LA = StridedLoader<T>;
  using Epi = StridedEpilogue<T>;
  LA la{A, a_mn, a_k, M, K};
  LA lb{B, b_mn, b_k, N, K};
  const bool vec_ok = (csC == 1) && (rsC % Cfg::V == 0) && (bsC % Cfg::V == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  Epi epi{C, rsC, csC, M, N, alpha, beta, vec_ok, bias_m, bias_n};
  const int a_kfast = iabs64(a_k) <= iabs64(a_mn);
  const int b_kfast = iabs64(b_k) <= iabs64(b_mn);
  dim3 grid((unsigned)ceil_div(N, Cfg::BN), (unsigned)ceil_div(M, Cfg::BM), (unsigned)(Batched ? batch : 1));
  if (grid.y > 65535) {  // fold very tall problems: launch in row bands
    const int64_t band = (int64_t)65535 * Cfg::BM;
    for (int64_t r = 0; r < M; r += band) {
      const int64_t mb = (M - r < band) ? M - r : band;
      int rc = launch_cfg<T, Cfg, Batched>(st, batch, mb, N, K, alpha, A + r * a_mn, a_mn, a_k, bsA, B, b_mn, b_k, bsB, beta,
                                           C + r * rsC, rsC, csC, bsC, bias_m ? bias_m + r : nullptr, bias_n);
      if (rc) return rc;
    }
    return AM_OK;
  }
  const int* wide_flag = nullptr;
  if constexpr (std::is_same<T, int64_t>::value && !Batched) {
    if (2.0 * (double)M * (double)N * (double)K >= 2.0e8) {
      // flags live in a small ring so back-to-back calls on different streams do not share one
      static std::atomic<unsigned> ring{0};
      void* base = nullptr;
      int rc = workspace(kWsMisc, 64 * sizeof(int) + 1024, &base);
      if (rc) return rc;
      int* flag = (int*)base + (ring++ % 64);
      AM_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st));
      i64_range_kernel<<<dim3((unsigned)(4 * sm_count()), 2), 256, 0, st>>>(A, a_mn, a_k, M, B, b_mn, b_k, N, K, flag);
      g_launch_count++;
      wide_flag = flag;
    }
  }
  int ma = stage_mode<T>(A, a_mn, a_k, Batched ? bsA : 0), mb = stage_mode<T>(B, b_mn, b_k, Batched ? bsB : 0);
  if (ma == kStageScalar || mb == kStageScalar) ma = mb = kStageScalar;     // mixed layouts: element-wise staging for both
#define AM_MODE(MA_, MB_) \
  if (ma == MA_ && mb == MB_) return launch_mode<T, Cfg, Batched, MA_, MB_>(st, grid, la, lb, epi, K, a_kfast, b_kfast, wide_flag, bsA, bsB, bsC);
  AM_MODE(kStageVecK, kStageVecMN)     // row-major A x row-major B
  AM_MODE(kStageVecMN, kStageVecK)     // column-major A x column-major B (and the C^T = B^T A^T form of the row-major case)
  AM_MODE(kStageVecK, kStageVecK)      // A x B^T
  AM_MODE(kStageVecMN, kStageVecMN)    // A^T x B
#undef AM_MODE
  return launch_mode<T, Cfg, Batched, kStageScalar, kStageScalar>(st, grid, la, lb, epi, K, a_kfast, b_kfast, wide_flag, bsA, bsB, bsC);
}

// batch == 0: plain GEMM.  batch >= 1: `batch` independent products, operand b at X + b*bsX (one launch, blockIdx.z = b).
template <class T>
static int gemm_simt_any(cudaStream_t st, int64_t batch, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t rsA,
                         int64_t csA, int64_t bsA, const T* B, int64_t rsB, int64_t csB, int64_t bsB, T beta, T* C,
                         int64_t rsC, int64_t csC, int64_t bsC, const T* bias_col = nullptr) {
  // Operand views in (mn_stride, k_stride) form.  A: mn = row, B: mn = column.
  int64_t a_mn = rsA, a_k = csA, b_mn = csB, b_k = rsB;
  const T *bias_m = nullptr, *bias_n = bias_col;     // bias is per column of C; the transposed form below makes it per row
  // Column-major-ish C (the CudaTensor default, data_structure.nim:44-58): compute
  // C^T = B^T A^T so the fast dimension of C maps onto the lanes' vector dimension.
  if (iabs64(rsC) < iabs64(csC)) {
    const T* tp = A; A = B; B = tp;
    int64_t t;
    t = a_mn; a_mn = b_mn; b_mn = t;
    t = a_k; a_k = b_k; b_k = t;
    t = M; M = N; N = t;
    t = rsC; rsC = csC; csC = t;
    t = bsA; bsA = bsB; bsB = t;
    bias_m = bias_col; bias_n = nullptr;
  }
  // pick the tile: full 128x128 tiles when they fill the chip, 64x64 otherwise
  const int sms = sm_count();
  const double nb = batch > 0 ? (double)batch : 1.0;
  auto eff = [&](int64_t bm, int64_t bn, int occ, double intrinsic) {
    const double tiles = (double)ceil_div(M, bm) * (double)ceil_div(N, bn) * nb;
    const double slots = (double)sms * occ;
    const double waves = std::ceil(tiles / slots);
    const double fill = ((double)M * (double)N * nb) / (tiles * (double)bm * (double)bn);
    return intrinsic * fill * tiles / (waves * slots);
  };
  const int occ_big = sizeof(T) == 8 ? 1 : 2;
  const bool big = eff(128, 128, occ_big, 1.0) >= eff(64, 64, 3, 0.75);
  if (batch > 0) {
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {           // grid.z limit
      const int64_t nbz = batch - b0 < 65535 ? batch - b0 : 65535;
      int rc = big ? launch_cfg<T, typename GemmCfgs<T>::Big, true>(st, nbz, M, N, K, alpha, A + b0 * bsA, a_mn, a_k, bsA, B + b0 * bsB,
                                                                    b_mn, b_k, bsB, beta, C + b0 * bsC, rsC, csC, bsC)
                   : launch_cfg<T, typename GemmCfgs<T>::Small, true>(st, nbz, M, N, K, alpha, A + b0 * bsA, a_mn, a_k, bsA, B + b0 * bsB,
                                                                      b_mn, b_k, bsB, beta, C + b0 * bsC, rsC, csC, bsC);
      if (rc) return rc;
    }
    return AM_OK;
  }
  if (big)
    return launch_cfg<T, typename GemmCfgs<T>::Big, false>(st, 1, M, N, K, alpha, A, a_mn, a_k, 0, B, b_mn, b_k, 0, beta, C,
                                                           rsC, csC, 0, bias_m, bias_n);
  return launch_cfg<T, typename GemmCfgs<T>::Small, false>(st, 1, M, N, K, alpha, A, a_mn, a_k, 0, B, b_mn, b_k, 0, beta, C,
                                                           rsC, csC, 0, bias_m, bias_n);
}

template <class T>
int gemm_simt(cudaStream_t st, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t rsA, int64_t csA,
              const T* B, int64_t rsB, int64_t csB, T beta, T* C, int64_t rsC, int64_t csC, const T* bias_col) {
  return gemm_simt_any<T>(st, 0, M, N, K, alpha, A, rsA, csA, 0, B, rsB, csB, 0, beta, C, rsC, csC, 0, bias_col);
}
template <class T>
int gemm_simt_batched(cudaStream_t st, int64_t batch, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t rsA,
                      int64_t csA, int64_t bsA, const T* B, int64_t rsB, int64_t csB, int64_t bsB, T beta, T* C,
                      int64_t rsC, int64_t csC, int64_t bsC) {
  if (batch <= 0) return AM_OK;
  return gemm_simt_any<T>(st, batch, M, N, K, alpha, A, rsA, csA, bsA, B, rsB, csB, bsB, beta, C, rsC, csC, bsC);
}

#define AM_INST_SIMT(T)                                                                                                  \
  template int gemm_simt<T>(cudaStream_t, int64_t, int64_t, int64_t, T, const T*, int64_t, int64_t, const T*, int64_t,   \
                            int64_t, T, T*, int64_t, int64_t, const T*);                                                 \
  template int gemm_simt_batched<T>(cudaStream_t, int64_t, int64_t, int64_t, int64_t, T, const T*, int64_t, int64_t,     \
                                    int64_t, const T*, int64_t, int64_t, int64_t, T, T*, int64_t, int64_t, int64_t);

}  // namespace am
