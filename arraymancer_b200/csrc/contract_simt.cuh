// Generic SIMT contraction core: C[m,n] (+)= sum_k A(m,k) * B(k,n) with pluggable operand
// loaders and epilogue.  One register-tiled, double-buffered shared-memory mainloop serves
//   * the integer GEMMs (IMAD, exact mod 2^n)                      -> gemm_simt.cu
//   * the f64 / exact-f32 strided GEMMs (DFMA / FFMA)               -> gemm_simt.cu
//   * the implicit-GEMM convolution fwd / dgrad / wgrad             -> conv.cu
// Operands are fetched straight from their (arbitrarily strided or gathered) global layout
// into registers, staged through padded shared memory, and consumed as 128-bit LDS vectors.
//
// Reference counterpart: the BLIS-style 5-loop nest + packing of
// laser/primitives/matrix_multiplication/gemm.nim:117-184 and gemm_packing.nim:24-99 —
// here the "packing" is the global->shared staging and the micro-kernel is the TMxTN
// register tile (gemm_ukernel_generator.nim:143-253).
#pragma once
#include "am_common.cuh"

namespace am {

template <class T, int BM_, int BN_, int BK_, int TM_, int TN_>
struct SimtCfg {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, TM = TM_, TN = TN_;
  static constexpr int V = 16 / (int)sizeof(T);            // elements per 128-bit vector
  static constexpr int TX = BN / TN, TY = BM / TM, NT = TX * TY;
  static constexpr int PAD = V;                             // keeps rows 16-B aligned, spreads banks
  static constexpr int LDA = BM + PAD, LDB = BN + PAD;
  static constexpr int EA = BM * BK / NT, EB = BN * BK / NT; // elements each thread stages per k-tile
  static_assert(TM % V == 0 && TN % V == 0, "register tile must be a multiple of the vector width");
  static_assert((BM * BK) % NT == 0 && (BN * BK) % NT == 0, "tile must divide evenly over threads");
  static_assert(NT % BM == 0 || BM % NT == 0, "mn-fast mapping needs NT | BM or BM | NT");
  static_assert(NT % BK == 0, "k-fast mapping needs BK | NT");
};

// ---------------------------------------------------------------------------- loaders
// Loader concept (L):
//   struct Slot;                                              per staged element, tile invariant
//   Slot slot(int64_t mn, int k_in_tile) const;               mn = global row (A) / column (B)
//   T    load(const Slot&, int64_t k_tile_base) const;        value at k = k_tile_base + k_in_tile
//   static constexpr int kMapping;   which tile dimension the lanes run along when staging:
//                                    0 = runtime flag, 1 = always mn, 2 = always k (compile-time
//                                    mappings let the index decode of gathers fold / CSE)
//
// Plain strided matrix: element (mn, k) at p[mn*mn_stride + k*k_stride], zero outside.
template <class T>
struct StridedLoader {
  static constexpr int kMapping = 0;
  const T* p;
  int64_t mn_stride, k_stride, MN, Kend;
  struct Slot { const T* ptr; int k; bool ok; };
  __device__ __forceinline__ Slot slot(int64_t mn, int k) const {
    return Slot{p + mn * mn_stride + (int64_t)k * k_stride, k, mn < MN};
  }
  __device__ __forceinline__ T load(const Slot& s, int64_t kt) const {
    return (s.ok && kt + s.k < Kend) ? s.ptr[kt * k_stride] : T(0);
  }
  // cp.async staging: source address + validity (invalid -> zero fill, the address is not dereferenced)
  static constexpr bool kAsync = true;
  __device__ __forceinline__ const T* addr(const Slot& s, int64_t kt, bool* ok) const {
    *ok = s.ok && kt + s.k < Kend;
    return *ok ? s.ptr + kt * k_stride : p;
  }
};

template <class L, class = void> struct LoaderIsAsync { static constexpr bool value = false; };
template <class L> struct LoaderIsAsync<L, typename std::enable_if<L::kAsync>::type> { static constexpr bool value = true; };

// global -> shared without a register round trip (LDGSTS); src_bytes = 0 zero-fills the destination
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(uint32_t dst_smem, const void* src, bool ok) {
  const int src_bytes = ok ? BYTES : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(dst_smem), "l"(src), "n"(BYTES), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------- epilogues
// Epilogue concept: void store(int64_t m, int64_t n0, const T (&v)[V], int z) — V consecutive
// columns n0..n0+V-1 of row m (bounds-checked inside); z = blockIdx.z (split-K slice).
template <class T>
struct StridedEpilogue {   // gemm_ukernel_generic.nim:96-125 semantics on a strided C
  T* C;
  int64_t rs, cs, M, N;
  T alpha, beta;
  bool vec_ok;             // cs == 1, 16-B aligned rows: whole vector in one transaction
  template <int V>
  __device__ __forceinline__ void store(int64_t m, int64_t n0, const T (&v)[V], int) const {
    if (m >= M || n0 >= N) return;
    T* row = C + m * rs;
    if (vec_ok && n0 + V <= N) {
      using Vec = typename std::conditional<sizeof(T) == 4, int4, longlong2>::type;
      union { Vec q; T e[V]; } o, c;
      if (beta != T(0)) c.q = *reinterpret_cast<const Vec*>(row + n0);
#pragma unroll
      for (int j = 0; j < V; j++) o.e[j] = epilogue_value<T>(alpha, v[j], beta, beta != T(0) ? c.e[j] : T(0));
      *reinterpret_cast<Vec*>(row + n0) = o.q;
    } else {
#pragma unroll
      for (int j = 0; j < V; j++) {
        if (n0 + j < N) {
          T* pc = row + (n0 + j) * cs;
          *pc = epilogue_value<T>(alpha, v[j], beta, beta != T(0) ? *pc : T(0));
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------- mainloop
// grid = (ceil(N/BN), ceil(M/BM), splits); slice z covers k in [z*k_per_split, min(K, (z+1)*k_per_split)).
template <class T, class Cfg, class LA, class LB, class Epi>
__global__ void __launch_bounds__(Cfg::NT)
contract_simt_kernel(const LA la, const LB lb, const Epi epi, int64_t K, int64_t k_per_split,
                     int a_kfast, int b_kfast, const int* __restrict__ wide_flag = nullptr) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, TM = Cfg::TM, TN = Cfg::TN, V = Cfg::V;
  constexpr int TX = Cfg::TX, TY = Cfg::TY, NT = Cfg::NT, LDA = Cfg::LDA, LDB = Cfg::LDB;
  constexpr int EA = Cfg::EA, EB = Cfg::EB;
  using Vec = typename std::conditional<sizeof(T) == 4, int4, longlong2>::type;

  __shared__ __align__(16) T As[2][BK][LDA];
  __shared__ __align__(16) T Bs[2][BK][LDB];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
  const int64_t kend = (kbeg + k_per_split < K) ? kbeg + k_per_split : K;

  // ---- staging slots: which (mn, k) of the tile each thread fetches, and where it lands
  const bool akf = LA::kMapping == 0 ? (a_kfast != 0) : (LA::kMapping == 2);
  const bool bkf = LB::kMapping == 0 ? (b_kfast != 0) : (LB::kMapping == 2);
  typename LA::Slot sa[EA];
  typename LB::Slot sb[EB];
  int oa[EA], ob[EB];
#pragma unroll
  for (int i = 0; i < EA; i++) {
    const int idx = tid + i * NT;
    const int mi = akf ? idx / BK : idx % BM;
    const int ki = akf ? idx % BK : idx / BM;
    sa[i] = la.slot(m0 + mi, ki);
    oa[i] = ki * LDA + mi;
  }
#pragma unroll
  for (int i = 0; i < EB; i++) {
    const int idx = tid + i * NT;
    const int ni = bkf ? idx / BK : idx % BN;
    const int ki = bkf ? idx % BK : idx / BN;
    sb[i] = lb.slot(n0 + ni, ki);
    ob[i] = ki * LDB + ni;
  }

  T acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = T(0);
  // float32 only: two-level accumulation with the reference's K blocking (kc = 512 for 4-byte types,
  // gemm_tiling.nim:310): fma-sequential inside a 512-deep block, blocks added in order — the same order
  // as laser's `C += AB_block` (gemm.nim:158-166), which keeps the error from growing like sqrt(K).
  constexpr bool kTwoLevel = std::is_same<T, float>::value;
  constexpr int kFlushTiles = 512 / BK;
  T acc2[kTwoLevel ? TM : 1][kTwoLevel ? TN : 1];
  if constexpr (kTwoLevel) {
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
      for (int j = 0; j < TN; j++) acc2[i][j] = T(0);
  }
  int flush_cnt = 0;

  const int64_t ntiles = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;
  // int64 only: a pre-pass (gemm_simt.cu::i64_range_kernel) found every operand element inside the
  // int32 range -> one IMAD.WIDE per multiply-accumulate instead of three IMADs, same bits.
  bool narrow = false;
  if constexpr (std::is_same<T, int64_t>::value) narrow = (wide_flag != nullptr) && (*wide_flag == 0);

  auto compute_tile = [&](int buf) {
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      union { Vec q[TM / V]; T e[TM]; } fa;
      union { Vec q[TN / V]; T e[TN]; } fb;
#pragma unroll
      for (int g = 0; g < TM / V; g++)
        fa.q[g] = *reinterpret_cast<const Vec*>(&As[buf][kk][g * (TY * V) + ty * V]);
#pragma unroll
      for (int g = 0; g < TN / V; g++)
        fb.q[g] = *reinterpret_cast<const Vec*>(&Bs[buf][kk][g * (TX * V) + tx * V]);
      if constexpr (std::is_same<T, int64_t>::value) {
        if (narrow) {
#pragma unroll
          for (int i = 0; i < TM; i++)
#pragma unroll
            for (int j = 0; j < TN; j++) acc[i][j] = mac_narrow_i64(fa.e[i], fb.e[j], acc[i][j]);
          continue;
        }
      }
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = mac<T>(fa.e[i], fb.e[j], acc[i][j]);
    }
    if constexpr (kTwoLevel) {
      if (++flush_cnt == kFlushTiles) {
        flush_cnt = 0;
#pragma unroll
        for (int i = 0; i < TM; i++)
#pragma unroll
          for (int j = 0; j < TN; j++) { acc2[i][j] = __fadd_rn(acc2[i][j], acc[i][j]); acc[i][j] = T(0); }
      }
    }
  };

  if constexpr (LoaderIsAsync<LA>::value && LoaderIsAsync<LB>::value) {
    // ---- cp.async (LDGSTS) staging: tile t+1 streams into the other buffer while tile t is consumed; no
    //      staging registers (the register-staged version had its loads sunk next to the STS by ptxas under the
    //      128-register cap: 14 % of all stall samples on that one STS, profiles/r01_bringup.md)
    const uint32_t as_base = (uint32_t)__cvta_generic_to_shared(&As[0][0][0]);
    const uint32_t bs_base = (uint32_t)__cvta_generic_to_shared(&Bs[0][0][0]);
    auto issue = [&](int buf, int64_t kt) {
#pragma unroll
      for (int i = 0; i < EA; i++) {
        bool ok;
        const T* src = la.addr(sa[i], kt, &ok);
        cp_async_zfill<(int)sizeof(T)>(as_base + (uint32_t)((buf * BK * LDA + oa[i]) * sizeof(T)), src, ok);
      }
#pragma unroll
      for (int i = 0; i < EB; i++) {
        bool ok;
        const T* src = lb.addr(sb[i], kt, &ok);
        cp_async_zfill<(int)sizeof(T)>(bs_base + (uint32_t)((buf * BK * LDB + ob[i]) * sizeof(T)), src, ok);
      }
      cp_async_commit();
    };
    if (ntiles > 0) issue(0, kbeg);
    for (int64_t t = 0; t < ntiles; t++) {
      cp_async_wait_all();
      __syncthreads();          // tile t visible to all; everyone is done reading the other buffer (tile t-1)
      if (t + 1 < ntiles) issue((int)((t + 1) & 1), kbeg + (t + 1) * BK);
      compute_tile((int)(t & 1));
    }
  } else {
    T ra[EA], rb[EB];
    if (ntiles > 0) {
#pragma unroll
      for (int i = 0; i < EA; i++) ra[i] = la.load(sa[i], kbeg);
#pragma unroll
      for (int i = 0; i < EB; i++) rb[i] = lb.load(sb[i], kbeg);
      T* a0 = &As[0][0][0];
      T* b0 = &Bs[0][0][0];
#pragma unroll
      for (int i = 0; i < EA; i++) a0[oa[i]] = ra[i];
#pragma unroll
      for (int i = 0; i < EB; i++) b0[ob[i]] = rb[i];
    }
    __syncthreads();
    for (int64_t t = 0; t < ntiles; t++) {
      const int buf = (int)(t & 1);
      const bool more = (t + 1 < ntiles);
      if (more) {      // global -> registers for the next tile while this one is consumed
        const int64_t kt = kbeg + (t + 1) * BK;
#pragma unroll
        for (int i = 0; i < EA; i++) ra[i] = la.load(sa[i], kt);
#pragma unroll
        for (int i = 0; i < EB; i++) rb[i] = lb.load(sb[i], kt);
      }
      compute_tile(buf);
      if (more) {
        T* a1 = &As[buf ^ 1][0][0];
        T* b1 = &Bs[buf ^ 1][0][0];
#pragma unroll
        for (int i = 0; i < EA; i++) a1[oa[i]] = ra[i];
#pragma unroll
        for (int i = 0; i < EB; i++) b1[ob[i]] = rb[i];
      }
      __syncthreads();
    }
  }

  if constexpr (kTwoLevel) {
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
      for (int j = 0; j < TN; j++) acc[i][j] = __fadd_rn(acc2[i][j], acc[i][j]);
  }

  // ---- epilogue: thread owns rows g*(TY*V)+ty*V+v, column groups h*(TX*V)+tx*V..+V
#pragma unroll
  for (int g = 0; g < TM / V; g++)
#pragma unroll
    for (int v = 0; v < V; v++) {
      const int64_t m = m0 + g * (TY * V) + ty * V + v;
#pragma unroll
      for (int h = 0; h < TN / V; h++) {
        T out[V];
#pragma unroll
        for (int j = 0; j < V; j++) out[j] = acc[g * V + v][h * V + j];
        epi.template store<V>(m, n0 + h * (TX * V) + tx * V, out, (int)blockIdx.z);
      }
    }
}

}  // namespace am
