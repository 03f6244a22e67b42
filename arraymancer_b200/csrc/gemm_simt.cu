// Strided SIMT GEMM for all four dtypes (int32/int64: IMAD, exact mod 2^n; f64: DFMA;
// f32: exact FFMA path used for small / skinny shapes and as the AM_F32_SIMT selector).
// Device-side replacement of laser gemm_strided (gemm.nim:192-273) for arbitrary
// (rowStride, colStride) views — negative, zero and transposed strides included.
#include "contract_simt.cuh"
#include "gemm_dispatch.h"

namespace am {

std::atomic<int64_t> g_launch_count{0};

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

// Pre-pass of the int64 GEMM: is every element of A and of B representable in int32?  (flag := 1 if not.)
// O(MK + KN) reads; lets the mainloop use one IMAD.WIDE per multiply-accumulate (bit-identical result).
__global__ void i64_range_kernel(const int64_t* __restrict__ A, int64_t a_mn, int64_t a_k, int64_t M,
                                 const int64_t* __restrict__ B, int64_t b_mn, int64_t b_k, int64_t N, int64_t K,
                                 int* __restrict__ wide_flag) {
  const bool isB = blockIdx.y == 1;
  const int64_t* X = isB ? B : A;
  const int64_t mn_stride = isB ? b_mn : a_mn, k_stride = isB ? b_k : a_k, MN = isB ? N : M;
  const bool k_fast = iabs64_dev(k_stride) <= iabs64_dev(mn_stride);
  const int64_t inner = k_fast ? K : MN, total = MN * K;
  bool wide = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = i / inner, in = i - o * inner;
    const int64_t v = k_fast ? X[o * mn_stride + in * k_stride] : X[in * mn_stride + o * k_stride];
    wide |= (v != (int64_t)(int32_t)v);
  }
  if (__any_sync(0xffffffffu, wide) && (threadIdx.x & 31) == 0) atomicOr(wide_flag, 1);
}

template <>
struct GemmCfgs<int32_t> {
  static constexpr int BK = 16;
  using Big = SimtCfg<int32_t, 128, 128, BK, 4, 8>;   // 512 threads: IMAD issues at half the FFMA rate, occupancy wins
  using Small = SimtCfg<int32_t, 64, 64, BK, 4, 4>;
};

template <class T, class Cfg>
static int launch_cfg(cudaStream_t st, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t a_mn,
                      int64_t a_k, const T* B, int64_t b_mn, int64_t b_k, T beta, T* C, int64_t rsC,
                      int64_t csC) {
  using LA = StridedLoader<T>;
  using Epi = StridedEpilogue<T>;
  LA la{A, a_mn, a_k, M, K};
  LA lb{B, b_mn, b_k, N, K};
  const bool vec_ok = (csC == 1) && (rsC % Cfg::V == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  Epi epi{C, rsC, csC, M, N, alpha, beta, vec_ok};
  const int a_kfast = iabs64(a_k) <= iabs64(a_mn);
  const int b_kfast = iabs64(b_k) <= iabs64(b_mn);
  dim3 grid((unsigned)ceil_div(N, Cfg::BN), (unsigned)ceil_div(M, Cfg::BM), 1);
  if (grid.y > 65535) {  // fold very tall problems: launch in row bands
    const int64_t band = (int64_t)65535 * Cfg::BM;
    for (int64_t r = 0; r < M; r += band) {
      const int64_t mb = (M - r < band) ? M - r : band;
      int rc = launch_cfg<T, Cfg>(st, mb, N, K, alpha, A + r * a_mn, a_mn, a_k, B, b_mn, b_k, beta,
                                  C + r * rsC, rsC, csC);
      if (rc) return rc;
    }
    return AM_OK;
  }
  const int* wide_flag = nullptr;
  if constexpr (std::is_same<T, int64_t>::value) {
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
  contract_simt_kernel<T, Cfg, LA, LA, Epi><<<grid, Cfg::NT, 0, st>>>(la, lb, epi, K, K, a_kfast, b_kfast, wide_flag);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  return AM_OK;
}

template <class T>
int gemm_simt(cudaStream_t st, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t rsA, int64_t csA,
              const T* B, int64_t rsB, int64_t csB, T beta, T* C, int64_t rsC, int64_t csC) {
  // Operand views in (mn_stride, k_stride) form.  A: mn = row, B: mn = column.
  int64_t a_mn = rsA, a_k = csA, b_mn = csB, b_k = rsB;
  // Column-major-ish C (the CudaTensor default, data_structure.nim:44-58): compute
  // C^T = B^T A^T so the fast dimension of C maps onto the lanes' vector dimension.
  if (iabs64(rsC) < iabs64(csC)) {
    const T* tp = A; A = B; B = tp;
    int64_t t;
    t = a_mn; a_mn = b_mn; b_mn = t;
    t = a_k; a_k = b_k; b_k = t;
    t = M; M = N; N = t;
    t = rsC; rsC = csC; csC = t;
  }
  // pick the tile: full 128x128 tiles when they fill the chip, 64x64 otherwise
  const int sms = sm_count();
  auto eff = [&](int64_t bm, int64_t bn, int occ, double intrinsic) {
    const double tiles = (double)ceil_div(M, bm) * (double)ceil_div(N, bn);
    const double slots = (double)sms * occ;
    const double waves = std::ceil(tiles / slots);
    const double fill = ((double)M * (double)N) / (tiles * (double)bm * (double)bn);
    return intrinsic * fill * tiles / (waves * slots);
  };
  const int occ_big = sizeof(T) == 8 ? 1 : 2;
  const bool big = eff(128, 128, occ_big, 1.0) >= eff(64, 64, 3, 0.75);
  if (big)
    return launch_cfg<T, typename GemmCfgs<T>::Big>(st, M, N, K, alpha, A, a_mn, a_k, B, b_mn, b_k, beta, C,
                                                    rsC, csC);
  return launch_cfg<T, typename GemmCfgs<T>::Small>(st, M, N, K, alpha, A, a_mn, a_k, B, b_mn, b_k, beta, C,
                                                    rsC, csC);
}

template int gemm_simt<float>(cudaStream_t, int64_t, int64_t, int64_t, float, const float*, int64_t, int64_t,
                              const float*, int64_t, int64_t, float, float*, int64_t, int64_t);
template int gemm_simt<double>(cudaStream_t, int64_t, int64_t, int64_t, double, const double*, int64_t, int64_t,
                               const double*, int64_t, int64_t, double, double*, int64_t, int64_t);
template int gemm_simt<int32_t>(cudaStream_t, int64_t, int64_t, int64_t, int32_t, const int32_t*, int64_t,
                                int64_t, const int32_t*, int64_t, int64_t, int32_t, int32_t*, int64_t, int64_t);
template int gemm_simt<int64_t>(cudaStream_t, int64_t, int64_t, int64_t, int64_t, const int64_t*, int64_t,
                                int64_t, const int64_t*, int64_t, int64_t, int64_t, int64_t*, int64_t, int64_t);

}  // namespace am
