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
  // kStageVecMN (mn_stride == 1): V consecutive mn starting at the slot's mn, one k.  The slot only knows whether its
  // FIRST mn is inside; the number of valid bytes is clamped against MN (the copy zero-fills the rest).
  __device__ __forceinline__ const T* vec_addr(const Slot& s, int64_t kt) const {
    return (s.ok && kt + s.k < Kend) ? s.ptr + kt * k_stride : p;
  }
  __device__ __forceinline__ int vec_bytes(const Slot& s, int64_t kt, int64_t) const {
    if (!s.ok || kt + s.k >= Kend) return 0;
    const int64_t mn = (s.ptr - p - (int64_t)s.k * k_stride);      // mn_stride == 1: element offset of the row start
    const int64_t left = MN - mn;
    return left >= (int64_t)(16 / sizeof(T)) ? 16 : (int)(left * (int64_t)sizeof(T));
  }
};

// 16-byte cp.async with a runtime source size: bytes beyond src_bytes are zero-filled (src_bytes == 0: no read at all)
__device__ __forceinline__ void cp_async_vec(uint32_t dst_smem, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}

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
  // fused bias of the linear layer (nnp_linear.nim:28-29 `result +.= bias`, a separate pass in the reference): added
  // AFTER the alpha / beta epilogue with its own rounding, exactly like the separate pass.  bias_n is indexed by the
  // column of this kernel's product, bias_m by its row (the C^T = B^T A^T form swaps them); both may be null.
  const T* bias_m = nullptr;
  const T* bias_n = nullptr;
  __device__ __forceinline__ T with_bias(T v, int64_t m, int64_t n) const {
    if (bias_m) v = add_nocontract<T>(v, bias_m[m]);
    if (bias_n) v = add_nocontract<T>(v, bias_n[n]);
    return v;
  }
  template <int V>
  __device__ __forceinline__ void store(int64_t m, int64_t n0, const T (&v)[V], int) const {
    if (m >= M || n0 >= N) return;
    T* row = C + m * rs;
    if (vec_ok && n0 + V <= N) {
      using Vec = typename std::conditional<sizeof(T) == 4, int4, longlong2>::type;
      union { Vec q; T e[V]; } o, c;
      if (beta != T(0)) c.q = *reinterpret_cast<const Vec*>(row + n0);
#pragma unroll
      for (int j = 0; j < V; j++) o.e[j] = with_bias(epilogue_value<T>(alpha, v[j], beta, beta != T(0) ? c.e[j] : T(0)), m, n0 + j);
      *reinterpret_cast<Vec*>(row + n0) = o.q;
    } else {
#pragma unroll
      for (int j = 0; j < V; j++) {
        if (n0 + j < N) {
          T* pc = row + (n0 + j) * cs;
          *pc = with_bias(epilogue_value<T>(alpha, v[j], beta, beta != T(0) ? *pc : T(0)), m, n0 + j);
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------- mainloop
// Staging modes of a StridedLoader operand (template parameters MA / MB):
//   kStageScalar  one cp.async (LDGSTS) of sizeof(T) per element — any stride, sign or alignment;
//   kStageVecMN   the operand's unit-stride dimension is mn (e.g. row-major B, column-major A): one 16-byte cp.async
//                 moves V consecutive mn of one k straight into the [k][mn] shared tile (partial tails zero-filled by
//                 the copy's src-size);
//   kStageVecK    the unit-stride dimension is k (row-major A, transposed B): one 128-bit LDG fetches V consecutive k
//                 of one mn into registers while the current tile is consumed, and V scalar STS transpose it into the
//                 [k][mn] tile afterwards.
// The vector modes need: unit stride along that dimension, the other stride a multiple of V, a 16-byte aligned base.
enum : int { kStageScalar = 0, kStageVecMN = 1, kStageVecK = 2 };

// grid = (ceil(N/BN), ceil(M/BM), splits); slice z covers k in [z*k_per_split, min(K, (z+1)*k_per_split)).
// Batched = true: blockIdx.z is a batch index instead (no K split): operand / result pointers advance by the batch
// strides bsA / bsB / bsC (StridedLoader / StridedEpilogue only).
template <class T, class Cfg, class LA, class LB, class Epi, bool Batched = false, int MA = 0, int MB = 0>
__global__ void __launch_bounds__(Cfg::NT)
contract_simt_kernel(const LA la_, const LB lb_, const Epi epi_, int64_t K, int64_t k_per_split,
                     int a_kfast, int b_kfast, const int* __restrict__ wide_flag = nullptr, int64_t bsA = 0,
                     int64_t bsB = 0, int64_t bsC = 0) {
  LA la = la_; LB lb = lb_; Epi epi = epi_;
  if constexpr (Batched) {
    la.p += (int64_t)blockIdx.z * bsA; lb.p += (int64_t)blockIdx.z * bsB; epi.C += (int64_t)blockIdx.z * bsC;
  }
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, TM = Cfg::TM, TN = Cfg::TN, V = Cfg::V;
  constexpr int TX = Cfg::TX, TY = Cfg::TY, NT = Cfg::NT, LDA = Cfg::LDA, LDB = Cfg::LDB;
  constexpr int EA = Cfg::EA, EB = Cfg::EB;
  using Vec = typename std::conditional<sizeof(T) == 4, int4, longlong2>::type;

  __shared__ __align__(16) T As[2][BK][LDA];
  __shared__ __align__(16) T Bs[2][BK][LDB];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t kbeg = Batched ? 0 : (int64_t)blockIdx.z * k_per_split;
  const int64_t kend = Batched ? K : ((kbeg + k_per_split < K) ? kbeg + k_per_split : K);

  // ---- staging slots: which (mn, k) of the tile each thread fetches, and where it lands
  const bool akf = LA::kMapping == 0 ? (a_kfast != 0) : (LA::kMapping == 2);
  const bool bkf = LB::kMapping == 0 ? (b_kfast != 0) : (LB::kMapping == 2);
  // (scalar mode: EA / EB element slots; vector modes: EA/V / EB/V slots of V elements each)
  constexpr int SA = MA == kStageScalar ? EA : (EA / V > 0 ? EA / V : 1);
  constexpr int SB = MB == kStageScalar ? EB : (EB / V > 0 ? EB / V : 1);
  static_assert(MA == kStageScalar || (EA % V == 0 && BK % V == 0 && BM % V == 0), "vector staging needs V | EA, BK, BM");
  static_assert(MB == kStageScalar || (EB % V == 0 && BK % V == 0 && BN % V == 0), "vector staging needs V | EB, BK, BN");
  typename LA::Slot sa[SA];
  typename LB::Slot sb[SB];
  int oa[SA], ob[SB];
#pragma unroll
  for (int i = 0; i < SA; i++) {
    const int idx = tid + i * NT;
    int mi, ki;
    if constexpr (MA == kStageVecMN) { mi = (idx % (BM / V)) * V; ki = idx / (BM / V); }
    else if constexpr (MA == kStageVecK) { ki = (idx % (BK / V)) * V; mi = idx / (BK / V); }
    else { mi = akf ? idx / BK : idx % BM; ki = akf ? idx % BK : idx / BM; }
    sa[i] = la.slot(m0 + mi, ki);
    oa[i] = ki * LDA + mi;
  }
#pragma unroll
  for (int i = 0; i < SB; i++) {
    const int idx = tid + i * NT;
    int ni, ki;
    if constexpr (MB == kStageVecMN) { ni = (idx % (BN / V)) * V; ki = idx / (BN / V); }
    else if constexpr (MB == kStageVecK) { ki = (idx % (BK / V)) * V; ni = idx / (BK / V); }
    else { ni = bkf ? idx / BK : idx % BN; ki = bkf ? idx % BK : idx / BN; }
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
    //      staging registers (the register-staged scalar version had its loads sunk next to the STS by ptxas under
    //      the 128-register cap: 14 % of all stall samples on that one STS, profiles/r01_bringup.md).  Operands in
    //      kStageVecK mode go through ONE 128-bit register per V elements instead (see the enum above).
    const uint32_t as_base = (uint32_t)__cvta_generic_to_shared(&As[0][0][0]);
    const uint32_t bs_base = (uint32_t)__cvta_generic_to_shared(&Bs[0][0][0]);
    Vec va[MA == kStageVecK ? SA : 1], vb[MB == kStageVecK ? SB : 1];
    // 128-bit load of V consecutive k of one mn; the (rare) vector that straddles Kend is assembled element-wise
    auto ldg_vec = [&](const auto& L, const auto& sl, int64_t kt) -> Vec {
      union { Vec q; T e[V]; } r;
      if (sl.ok && kt + sl.k + V <= L.Kend) {
        r.q = __ldg(reinterpret_cast<const Vec*>(sl.ptr + kt));          // k_stride == 1 in this mode
      } else {
#pragma unroll
        for (int v = 0; v < V; v++) r.e[v] = (sl.ok && kt + sl.k + v < L.Kend) ? sl.ptr[kt + v] : T(0);
      }
      return r.q;
    };
    auto issue = [&](int buf, int64_t kt) {
      if constexpr (MA == kStageScalar) {
#pragma unroll
        for (int i = 0; i < SA; i++) {
          bool ok;
          const T* src = la.addr(sa[i], kt, &ok);
          cp_async_zfill<(int)sizeof(T)>(as_base + (uint32_t)((buf * BK * LDA + oa[i]) * sizeof(T)), src, ok);
        }
      } else if constexpr (MA == kStageVecMN) {
#pragma unroll
        for (int i = 0; i < SA; i++)
          cp_async_vec(as_base + (uint32_t)((buf * BK * LDA + oa[i]) * sizeof(T)), la.vec_addr(sa[i], kt), la.vec_bytes(sa[i], kt, m0));
      } else {
#pragma unroll
        for (int i = 0; i < SA; i++) va[i] = ldg_vec(la, sa[i], kt);
      }
      if constexpr (MB == kStageScalar) {
#pragma unroll
        for (int i = 0; i < SB; i++) {
          bool ok;
          const T* src = lb.addr(sb[i], kt, &ok);
          cp_async_zfill<(int)sizeof(T)>(bs_base + (uint32_t)((buf * BK * LDB + ob[i]) * sizeof(T)), src, ok);
        }
      } else if constexpr (MB == kStageVecMN) {
#pragma unroll
        for (int i = 0; i < SB; i++)
          cp_async_vec(bs_base + (uint32_t)((buf * BK * LDB + ob[i]) * sizeof(T)), lb.vec_addr(sb[i], kt), lb.vec_bytes(sb[i], kt, n0));
      } else {
#pragma unroll
        for (int i = 0; i < SB; i++) vb[i] = ldg_vec(lb, sb[i], kt);
      }
      cp_async_commit();
    };
    // registers of the kStageVecK operands -> [k][mn] tile (transposing scalar stores)
    auto park = [&](int buf) {
      if constexpr (MA == kStageVecK) {
        T* a1 = &As[buf][0][0];
#pragma unroll
        for (int i = 0; i < SA; i++) {
          union { Vec q; T e[V]; } r; r.q = va[i];
#pragma unroll
          for (int v = 0; v < V; v++) a1[oa[i] + v * LDA] = r.e[v];
        }
      }
      if constexpr (MB == kStageVecK) {
        T* b1 = &Bs[buf][0][0];
#pragma unroll
        for (int i = 0; i < SB; i++) {
          union { Vec q; T e[V]; } r; r.q = vb[i];
#pragma unroll
          for (int v = 0; v < V; v++) b1[ob[i] + v * LDB] = r.e[v];
        }
      }
    };
    if (ntiles > 0) { issue(0, kbeg); park(0); }
    for (int64_t t = 0; t < ntiles; t++) {
      cp_async_wait_all();
      __syncthreads();          // tile t visible to all; everyone is done reading the other buffer (tile t-1)
      const bool more = t + 1 < ntiles;
      if (more) issue((int)((t + 1) & 1), kbeg + (t + 1) * BK);
      compute_tile((int)(t & 1));
      if (more) park((int)((t + 1) & 1));
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
