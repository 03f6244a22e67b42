// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (C++17 + OpenMP) of Arraymancer's laser `gemm_strided`
// (BLIS/Goto 5-loop GEMM with arbitrary strides).  Only tests/, bench.py's
// cpu_baseline / --impl reference leg and __graft_entry__.smoke() may use it,
// and only as the checker / reported CPU baseline.
//
// The reference itself (Nim) cannot be compiled in this image (no nim/nimble,
// third-party Nim deps not vendored) so this file restates the algorithm; it is
// pinned against every literal known-answer vector the reference's own tests hold
// for this path (tests/golden/*.json, tests/test_oracle_golden.py).
//
// Reference files followed (relative to /root/reference/src/arraymancer/laser/
// primitives/matrix_multiplication/):
//   gemm.nim:57-109      gebp_mkernel      -> gebp_mkernel()
//   gemm.nim:117-184     gemm_impl         -> gemm_impl()
//   gemm.nim:192-273     gemm_strided      -> gemm_strided()
//   gemm_tiling.nim:276-348  partitionMNK/newTiles -> Tiles
//   gemm_packing.nim:24-99   pack_A_mc_kc / pack_B_kc_nc
//   gemm_ukernel_generator.nim:143-253  register-tile FMA loop -> ukernel()
//   gemm_ukernel_generic.nim:53-76,96-125 epilogues -> epilogue_full/epilogue_edge
//   gemm_utils.nim:36-60     MatrixView
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace laser_oracle {

// gemm_utils.nim:36-60 — (ptr, rowStride, colStride) view, element (r,c) at r*rs + c*cs.
template <class T>
struct MatrixView {
  T* buffer;
  int64_t rs, cs;
  T& at(int64_t r, int64_t c) const { return buffer[r * rs + c * cs]; }
  MatrixView stride(int64_t r, int64_t c) const { return {buffer + r * rs + c * cs, rs, cs}; }
};

// Integer arithmetic wraps mod 2^n (gemm_ukernel_avx2.nim:15-16 mullo+add; the
// published benchmark is built -d:danger so the scalar paths wrap too): do it in
// the unsigned type to keep it defined behaviour in C++.
template <class T> struct arith {
  static T mul(T a, T b) { return a * b; }
  static T add(T a, T b) { return a + b; }
};
template <> struct arith<int32_t> {
  static int32_t mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
  static int32_t add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
};
template <> struct arith<int64_t> {
  static int64_t mul(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }
  static int64_t add(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
};

// One multiply-accumulate of the hot loop.  FMA ISAs (AVX_FMA/AVX512:
// gemm_ukernel_avx_fma.nim:27,42) fuse; SSE/AVX (gemm_ukernel_sse.nim:15-16,
// gemm_ukernel_avx.nim:15-19) round twice.  This TU is compiled with
// -ffp-contract=off so the choice below is exactly what runs.
template <class T, bool FMA>
static inline T mac(T a, T b, T acc) {
  if constexpr (std::is_floating_point<T>::value) {
    if constexpr (FMA) return std::fma(a, b, acc);
    else return acc + a * b;
  } else {
    return arith<T>::add(acc, arith<T>::mul(a, b));
  }
}

static inline int64_t round_step_down(int64_t x, int64_t step) { return x - x % step; }
static inline int64_t round_step_up(int64_t x, int64_t step) { return ((x + step - 1) / step) * step; }

// gemm_tiling.nim:251-348
template <class T>
struct Tiles {
  T* a = nullptr;
  T* b = nullptr;
  int64_t mc, nc, kc;
  int64_t ic_num_tasks, upanelA_size;
  Tiles(int64_t M, int64_t N, int64_t K, int MR, int NR) {
    nc = N;                                                   // gemm_tiling.nim:282
    mc = std::min<int64_t>(768 / (int64_t)sizeof(T), M);      // :309
    kc = std::min<int64_t>(2048 / (int64_t)sizeof(T), K);     // :310
    ic_num_tasks = mc > 0 ? (M + mc - 1) / mc : 0;            // :331
    upanelA_size = kc * round_step_up(mc, MR);                // :335
    size_t bufA = sizeof(T) * (size_t)upanelA_size * (size_t)ic_num_tasks;
    size_t bufB = sizeof(T) * (size_t)kc * (size_t)round_step_up(nc, NR);
    a = (T*)aligned_alloc(64, ((bufA + 63) / 64 + 1) * 64);
    b = (T*)aligned_alloc(64, ((bufB + 63) / 64 + 1) * 64);
  }
  ~Tiles() { free(a); free(b); }
  Tiles(const Tiles&) = delete;
};

// gemm_packing.nim:24-55 — A[mc,kc] strided -> MR-tall micro-panels, k-major, zero padded.
template <class T, int MR>
static void pack_A_mc_kc(T* __restrict buffer, int64_t mc, int64_t kc, MatrixView<const T> A) {
  const int64_t unroll_stop = round_step_down(mc, MR);
  for (int64_t i = 0; i < unroll_stop; i += MR)
    for (int64_t k = 0; k < kc; k++)
      for (int ii = 0; ii < MR; ii++)
        buffer[i * kc + k * MR + ii] = A.buffer[(i + ii) * A.rs + k * A.cs];
  const int64_t remainder = mc - unroll_stop;
  if (remainder > 0) {
    T* off = buffer + kc * unroll_stop;
    for (int64_t k = 0; k < kc; k++) {
      for (int64_t i = 0; i < remainder; i++) off[k * MR + i] = A.at(unroll_stop + i, k);
      for (int64_t i = remainder; i < MR; i++) off[k * MR + i] = T(0);
    }
  }
}

// gemm_packing.nim:63-99 — B[kc,nc] strided -> NR-wide micro-panels, zero padded;
// the full-panel loop is its own `omp parallel for` (:82).
template <class T, int NR>
static void pack_B_kc_nc(T* __restrict buffer, int64_t kc, int64_t nc, MatrixView<const T> B) {
  const int64_t unroll_stop = round_step_down(nc, NR);
#pragma omp parallel for
  for (int64_t j = 0; j < unroll_stop; j += NR)
    for (int64_t k = 0; k < kc; k++)
      for (int jj = 0; jj < NR; jj++)
        buffer[j * kc + k * NR + jj] = B.buffer[k * B.rs + (j + jj) * B.cs];
  const int64_t remainder = nc - unroll_stop;
  if (remainder > 0) {
    T* off = buffer + kc * unroll_stop;
    for (int64_t k = 0; k < kc; k++) {
      for (int64_t j = 0; j < remainder; j++) off[k * NR + j] = B.at(k, unroll_stop + j);
      for (int64_t j = remainder; j < NR; j++) off[k * NR + j] = T(0);
    }
  }
}

// gemm_ukernel_generator.nim:143-253 — AB[MR][NR] = sum_k A[k*MR+i]*B[k*NR+j], one
// accumulator per (i,j), k strictly ascending, no reassociation.
template <class T, int MR, int NR, bool FMA>
static inline void ukernel(int64_t kc, const T* __restrict A, const T* __restrict B, T (&AB)[MR][NR]) {
  for (int i = 0; i < MR; i++)
    for (int j = 0; j < NR; j++) AB[i][j] = T(0);
  for (int64_t k = 0; k < kc; k++) {
    const T* a = A + k * MR;
    const T* b = B + k * NR;
    for (int i = 0; i < MR; i++) {
      const T ai = a[i];
#pragma omp simd
      for (int j = 0; j < NR; j++) AB[i][j] = mac<T, FMA>(ai, b[j], AB[i][j]);
    }
  }
}

// gemm_ukernel_generic.nim:53-76 — full tile epilogue.
template <class T, int MR, int NR>
static inline void epilogue_full(T alpha, const T (&AB)[MR][NR], T beta, MatrixView<T> vC) {
  if (beta == T(0)) {
    for (int i = 0; i < MR; i++)
      for (int j = 0; j < NR; j++) vC.at(i, j) = T(0);
  } else if (beta != T(1)) {
    for (int i = 0; i < MR; i++)
      for (int j = 0; j < NR; j++) vC.at(i, j) = arith<T>::mul(vC.at(i, j), beta);
  }
  if (alpha == T(1)) {
    for (int i = 0; i < MR; i++)
      for (int j = 0; j < NR; j++) vC.at(i, j) = arith<T>::add(vC.at(i, j), AB[i][j]);
  } else {
    for (int i = 0; i < MR; i++)
      for (int j = 0; j < NR; j++)
        vC.at(i, j) = arith<T>::add(vC.at(i, j), arith<T>::mul(alpha, AB[i][j]));
  }
}

// gemm_ukernel_generic.nim:96-125 — edge tile epilogue (mr x nr valid).
template <class T, int MR, int NR>
static inline void epilogue_edge(T alpha, const T (&AB)[MR][NR], T beta, MatrixView<T> vC, int64_t mr,
                                 int64_t nr) {
  if (beta == T(0)) {
    if (alpha == T(1)) {
      for (int64_t i = 0; i < mr; i++)
        for (int64_t j = 0; j < nr; j++) vC.at(i, j) = AB[i][j];
    } else {
      for (int64_t i = 0; i < mr; i++)
        for (int64_t j = 0; j < nr; j++) vC.at(i, j) = arith<T>::mul(alpha, AB[i][j]);
    }
  } else {
    for (int64_t i = 0; i < mr; i++)
      for (int64_t j = 0; j < nr; j++) vC.at(i, j) = arith<T>::mul(vC.at(i, j), beta);
    if (alpha == T(1)) {
      for (int64_t i = 0; i < mr; i++)
        for (int64_t j = 0; j < nr; j++) vC.at(i, j) = arith<T>::add(vC.at(i, j), AB[i][j]);
    } else {
      for (int64_t i = 0; i < mr; i++)
        for (int64_t j = 0; j < nr; j++)
          vC.at(i, j) = arith<T>::add(vC.at(i, j), arith<T>::mul(alpha, AB[i][j]));
    }
  }
}

// gemm.nim:57-109 — macro kernel: jr (step NR) x ir (step MR).  Like the reference
// (`for jr in ||(0, nc-1, NR, "taskloop")`, gemm.nim:82) the jr loop is an OpenMP taskloop:
// the tasks a thread creates for its ic tile are also run by team threads that have no ic
// tile left (they wait at the end of the parallel region, a task scheduling point), so a
// problem with fewer ic tiles than threads still uses the whole team.  Every (ir,jr) tile
// is independent, the result does not depend on the schedule.
template <class T, int MR, int NR, bool FMA>
static void gebp_mkernel(int64_t mc, int64_t nc, int64_t kc, T alpha, const T* packA, const T* packB, T beta,
                         MatrixView<T> mcncC) {
#pragma omp taskloop default(shared)
  for (int64_t jr = 0; jr < nc; jr += NR) {
    const int64_t nr = std::min<int64_t>(nc - jr, NR);
    for (int64_t ir = 0; ir < mc; ir += MR) {
      const int64_t mr = std::min<int64_t>(mc - ir, MR);
      MatrixView<T> c_aux = mcncC.stride(ir, jr);
      alignas(64) T AB[MR][NR];
      ukernel<T, MR, NR, FMA>(kc, packA + ir * kc, packB + jr * kc, AB);
      if (nr == NR && mr == MR) epilogue_full<T, MR, NR>(alpha, AB, beta, c_aux);
      else epilogue_edge<T, MR, NR>(alpha, AB, beta, c_aux, mr, nr);
    }
  }
}

// gemm.nim:117-184
template <class T, int MR, int NR, bool FMA>
static void gemm_impl(int64_t M, int64_t N, int64_t K, T alpha, MatrixView<const T> vA, MatrixView<const T> vB,
                      T beta, MatrixView<T> vC, Tiles<T>& tiles, int max_threads) {
  const bool parallelize = (double)M * (double)N * (double)K > 128.0 * 128.0 * 128.0;  // pt = 128
  const int64_t nc = N;
  (void)max_threads;
  for (int64_t pc = 0; pc < K; pc += tiles.kc) {           // serial, ascending
    const int64_t kc = std::min<int64_t>(K - pc, tiles.kc);
    pack_B_kc_nc<T, NR>(tiles.b, kc, nc, vB.stride(pc, 0));
    const T beta_pc = (pc == 0) ? beta : T(1);             // gemm.nim:166
#pragma omp parallel if (parallelize)                      // gemm.nim:168 omp_parallel_if(parallelize)
    {
#pragma omp for nowait                                     // gemm.nim:171 `omp for nowait` over the ic tiles
      for (int64_t icb = 0; icb < tiles.ic_num_tasks; icb++) {
        T* packA = tiles.a + icb * tiles.upanelA_size;
        const int64_t ic = icb * tiles.mc;
        const int64_t mc = std::min<int64_t>(M - ic, tiles.mc);
        pack_A_mc_kc<T, MR>(packA, mc, kc, vA.stride(ic, pc));
        gebp_mkernel<T, MR, NR, FMA>(mc, nc, kc, alpha, packA, tiles.b, beta_pc, vC.stride(ic, 0));
      }
    }
  }
}

// ISA variants of the reference's dispatch (gemm.nim:237-272, gemm_tiling.nim:89-219).
enum Variant : int {
  kDefaultBuild = 0,   // AVX+FMA f32 6x16 / f64 6x8, AVX2 i32 6x16, SSE2 i64 6x4 (what ran the published 0.14 s)
  kAvx512 = 1,         // -d:avx512: f32/i32 14x32, f64/i64 14x16
  kGeneric = 2,        // 2x1 scalar, no FMA contraction
  kSseNoFma = 3,       // SSE/AVX float kernels: separate mul+add (two roundings)
};

template <class T>
static void gemm_strided(int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t rsA, int64_t csA,
                         const T* B, int64_t rsB, int64_t csB, T beta, T* C, int64_t rsC, int64_t csC,
                         int variant, int threads) {
  if (M <= 0 || N <= 0) return;
  // K == 0: the pc loop never runs, C untouched even if beta != 1 (gemm.nim:203 TODO).
  MatrixView<const T> vA{A, rsA, csA}, vB{B, rsB, csB};
  MatrixView<T> vC{C, rsC, csC};
#ifdef _OPENMP
  int prev = omp_get_max_threads();
  if (threads > 0) omp_set_num_threads(threads);
#endif
#define LASER_RUN(MR_, NR_, FMA_)                                                              \
  do {                                                                                         \
    Tiles<T> tiles(M, N, K, MR_, NR_);                                                         \
    gemm_impl<T, MR_, NR_, FMA_>(M, N, K, alpha, vA, vB, beta, vC, tiles, threads);            \
  } while (0)
  constexpr bool is32 = sizeof(T) == 4;
  constexpr bool isfp = std::is_floating_point<T>::value;
  switch (variant) {
    case kAvx512:
      if constexpr (is32) LASER_RUN(14, 32, true); else LASER_RUN(14, 16, true);
      break;
    case kGeneric:
      LASER_RUN(2, 1, false);
      break;
    case kSseNoFma:
      if constexpr (is32) LASER_RUN(6, 8, false); else LASER_RUN(6, 4, false);
      break;
    default:
      if constexpr (isfp) { if constexpr (is32) LASER_RUN(6, 16, true); else LASER_RUN(6, 8, true); }
      else { if constexpr (is32) LASER_RUN(6, 16, true); else LASER_RUN(6, 4, true); }
  }
#undef LASER_RUN
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(prev);
#endif
}

}  // namespace laser_oracle
