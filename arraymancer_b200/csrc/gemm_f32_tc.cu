// float32 GEMM on the 5th-gen tensor cores with a 3xTF32 split that keeps fp32 accuracy.
//
//   C = alpha * A*B + beta*C,  A = Ahi + Alo (+ ~2^-22 |A|),  B = Bhi + Blo
//   A*B ~= Alo*Bhi + Ahi*Blo + Ahi*Bhi   — three kind::tf32 UMMAs accumulating in fp32 in TMEM.
//
// Two kernels:
//  1. split_pack_kernel — the "packing" pass (role of laser's pack_A_mc_kc / pack_B_kc_nc,
//     gemm_packing.nim:24-99): reads an operand through ANY (row, col) stride pair — negative,
//     zero, transposed — and writes two K-major planes (hi, lo), tf32-rounded, zero padded to
//     tile multiples.  O(MK + KN) bytes, ~3 % of a 16384^3 GEMM.
//  2. gemm_tf32x3_kernel<CG, BK> — warp-specialised mainloop: warp 0 = TMA producer
//     (cp.async.bulk.tensor, 128B/64B-swizzled K-major tiles, mbarrier ring), warp 1 = UMMA issuer
//     (one elected lane, tcgen05.mma kind::tf32, accumulators in TMEM), warps 2-5 = epilogue
//     (tcgen05.ld -> alpha/beta -> strided global stores; lanes run along C's unit-stride dim).
//     CG = 2: CTA pair (cta_group::2), 256x512 output tile per pair = all 512 TMEM columns of both
//     SMs; each CTA stages its 128 rows of A and half of B, halving smem/L2 operand traffic.
//     CG = 1: single CTA, 128x256 tile (bring-up / comparison).
//
// Replaces the cuBLAS call behind CudaTensor `*` (tensor/backend/cublas.nim:142-170) and the
// float branch of gemm (tensor/operators_blas_l2l3.nim:58-71) for large shapes.
#include <cuda.h>
#include <cstdlib>

#include "am_common.cuh"
#include "gemm_dispatch.h"
#include "ptx_sm100.cuh"

namespace am {

// ------------------------------------------------------------------ split / pack pre-pass
// out planes: [Rpad][Kpad] row-major (K contiguous).  (r, k) of X at X[r*r_stride + k*k_stride].
__global__ void __launch_bounds__(256)
split_pack_kernel(const float* __restrict__ X, int64_t R, int64_t K, int64_t r_stride, int64_t k_stride,
                  float* __restrict__ hi, float* __restrict__ lo, int64_t Kpad, int k_fast) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int64_t k0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  float v[4];
  if (k_fast) {
    const int64_t k = k0 + tx;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int64_t r = r0 + ty + 8 * i;
      v[i] = (r < R && k < K) ? X[r * r_stride + k * k_stride] : 0.f;
    }
  } else {   // rows are the fast source dimension: read along r, transpose through smem
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int64_t r = r0 + tx, k = k0 + ty + 8 * i;
      tile[ty + 8 * i][tx] = (r < R && k < K) ? X[r * r_stride + k * k_stride] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = tile[tx][ty + 8 * i];
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int64_t r = r0 + ty + 8 * i;
    const float x = v[i];
    float h, l;
    if (isfinite(x)) { h = ptx::to_tf32_rna(x); l = ptx::to_tf32_rna(x - h); if (!isfinite(h)) { h = x; l = 0.f; } }
    else { h = x; l = 0.f; }
    const int64_t o = r * Kpad + k0 + tx;
    hi[o] = h;
    lo[o] = l;
  }
}

// Vector variant for the common case (unit stride along one source dimension, the other stride and the extents
// multiples of 4, 16-byte aligned base): 64 x 64 tiles, 128-bit loads along the contiguous source dimension and
// 128-bit stores along k.  The scalar kernel above moved 2.2 TB/s on a 32768^2 operand (5.8 ms of every sharded step).
__device__ __forceinline__ void split4(const float4 x, float4* h, float4* l) {
  const float xs[4] = {x.x, x.y, x.z, x.w};
  float hs[4], ls[4];
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const float v = xs[e];
    float hh = v, ll = 0.f;
    if (isfinite(v)) { hh = ptx::to_tf32_rna(v); ll = ptx::to_tf32_rna(v - hh); if (!isfinite(hh)) { hh = v; ll = 0.f; } }
    hs[e] = hh; ls[e] = ll;
  }
  *h = make_float4(hs[0], hs[1], hs[2], hs[3]);
  *l = make_float4(ls[0], ls[1], ls[2], ls[3]);
}

__global__ void __launch_bounds__(256)
split_pack_vec_kernel(const float* __restrict__ X, int64_t R, int64_t K, int64_t r_stride, int64_t k_stride,
                      float* __restrict__ hi, float* __restrict__ lo, int64_t Kpad, int k_fast) {
  __shared__ float tile[64][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16
  const int64_t k0 = (int64_t)blockIdx.x * 64, r0 = (int64_t)blockIdx.y * 64;
  if (k_fast) {                                               // k contiguous in the source: straight through
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int64_t r = r0 + ty + 16 * i, k = k0 + 4 * tx;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < R && k < K) x = __ldg(reinterpret_cast<const float4*>(X + r * r_stride + k));
      float4 h, l;
      split4(x, &h, &l);
      if (k < Kpad) {
        *reinterpret_cast<float4*>(hi + r * Kpad + k) = h;
        *reinterpret_cast<float4*>(lo + r * Kpad + k) = l;
      }
    }
    return;
  }
  // rows contiguous in the source: read 128-bit pieces along r, transpose through shared memory
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int64_t r = r0 + 4 * tx, k = k0 + ty + 16 * i;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < R && k < K) x = __ldg(reinterpret_cast<const float4*>(X + k * k_stride + r));
    tile[ty + 16 * i][4 * tx + 0] = x.x; tile[ty + 16 * i][4 * tx + 1] = x.y;
    tile[ty + 16 * i][4 * tx + 2] = x.z; tile[ty + 16 * i][4 * tx + 3] = x.w;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int rl = ty + 16 * i;
    const int64_t r = r0 + rl, k = k0 + 4 * tx;
    const float4 x = make_float4(tile[4 * tx + 0][rl], tile[4 * tx + 1][rl], tile[4 * tx + 2][rl], tile[4 * tx + 3][rl]);
    float4 h, l;
    split4(x, &h, &l);
    if (k < Kpad) {
      *reinterpret_cast<float4*>(hi + r * Kpad + k) = h;
      *reinterpret_cast<float4*>(lo + r * Kpad + k) = l;
    }
  }
}

// ------------------------------------------------------------------ mainloop
// Why the accumulators live in REGISTERS, not only in TMEM: the tensor core adds each MMA result
// into its fp32 accumulator with truncation, a bias that grows linearly with the length of the
// accumulation chain (measured on B200: rel. error ~7e-9 * K for one chain over all of K, i.e.
// 1e-4 at K = 16384 — not fp32 accuracy; profiles/r01_bringup.md).  So a chain only spans
// `flush_kb` k-blocks (64-128 k): each chain starts with accumulate = 0 into one of two 256-column
// TMEM buffers, and the 8 accumulate warps drain the finished buffer with tcgen05.ld and add it
// into their register tile with round-to-nearest FADDs while the next chain fills the other buffer.
template <int CG>
struct TcCfg {
  static constexpr int BK = 32;                      // floats per k-block = one 128-B swizzle row
  static constexpr int ROWS_A = 128;                 // A rows staged per CTA (= UMMA M per CTA)
  static constexpr int ROWS_B = (CG == 2) ? 128 : 256;   // B rows staged per CTA (pair: half of the 256 columns)
  static constexpr int TILE_M = 128 * CG;            // output tile of the CTA (pair)
  static constexpr int TILE_N = 256;
  static constexpr int TMEM_COLS = 512;              // two ping-pong chain buffers of 256 columns
  static constexpr int ROW_BYTES = BK * 4;
  static constexpr int SBO = 8 * ROW_BYTES;
  static constexpr int BOX_BYTES = 128 * ROW_BYTES;  // one TMA box: 128 rows x 32 floats = 16 KB
  static constexpr int A_BYTES = BOX_BYTES;          // per plane
  static constexpr int B_BYTES = (ROWS_B / 128) * BOX_BYTES;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;          // 64 KB (pair) / 96 KB (single)
  static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;              // 3 (pair) / 2 (single)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int NTHREADS = 384;               // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: accumulate
};

struct TcArgs {
  int64_t M, N;            // logical output extent (rows = TMEM lanes, cols = TMEM columns)
  int kblocks;             // Kpad / 32
  int flush_kb;            // k-blocks per tensor-core accumulation chain
  int tiles_m, tiles_n;
  int group;               // M-tiles per rasterisation group
  unsigned* sync_counter;  // grid barrier of the persistent schedule (zeroed before the launch); null = no barrier
  float* C; int64_t rsC, csC;
  float alpha, beta;
  // fused all-gather: when npeers > 0 the epilogue stores every result element to the same offset of each peer's
  // copy of C (peer[] holds device pointers mapped over NVLink, this GPU's own copy included) instead of to C
  int npeers;
  int self;                // which peer[] entry is this GPU's own copy
  float* peer[8];
  // fused bias of the linear layer (added after the alpha / beta epilogue, separately rounded like the reference's
  // `result +.= bias` pass, nnp_linear.nim:28-29): indexed by the TMEM lane (row m of this kernel) or by its column
  const float* bias_lane;
  const float* bias_col;
};

// tile index -> (tm, tn): groups of |group| tiles of one dimension, that dimension fastest inside a group, so that
// any window of consecutive tile indices (= the tiles resident at the same time) is a compact 2-D block
__device__ __forceinline__ void tile_coords(const TcArgs& p, int tile, int* tm, int* tn) {
  if (p.group > 0) {
    const int GROUP = p.group, group_size = GROUP * p.tiles_n, g = tile / group_size, first_m = g * GROUP;
    const int gm = (p.tiles_m - first_m < GROUP) ? (p.tiles_m - first_m) : GROUP;
    *tm = first_m + (tile % group_size) % gm;
    *tn = (tile % group_size) / gm;
  } else {
    const int GROUP = -p.group, group_size = GROUP * p.tiles_m, g = tile / group_size, first_n = g * GROUP;
    const int gn = (p.tiles_n - first_n < GROUP) ? (p.tiles_n - first_n) : GROUP;
    *tn = first_n + (tile % group_size) % gn;
    *tm = (tile % group_size) / gn;
  }
}

// PERSISTENT kernel: one CTA (pair) per SM (pair), tiles strided over the clusters.  All pipelines (smem ring, TMEM
// chain buffers) keep running across tile boundaries, so the epilogue of tile i overlaps the mainloop of tile i+1.
// Before starting the loads of its next tile each producer passes a grid-wide barrier (global counter, bounded spin):
// tiles of one wave then march through K in lockstep, which is what lets the 126 MB L2 capture the sharing of A/B
// panels between co-resident tiles — without it the tiles drift apart over the waves and most panel reads go to
// DRAM (measured 718-1221 GB per 32768^3 launch vs ~250 GB ideal; the extra DRAM power costs clocks under the cap).
template <int CG>
__global__ void __launch_bounds__(384, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                   const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                   const TcArgs p) {
  using Cfg = TcCfg<CG>;
  constexpr int BK = Cfg::BK;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;      // SW128 needs 1024-B alignment
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
  auto stage_base = [&](int s) { return smem_base + (uint32_t)s * Cfg::STAGE_BYTES; };
  // stage layout: [A_hi | A_lo | B_hi | B_lo]
  constexpr uint32_t OFF_AHI = 0, OFF_ALO = Cfg::A_BYTES, OFF_BHI = 2 * Cfg::A_BYTES, OFF_BLO = 2 * Cfg::A_BYTES + Cfg::B_BYTES;

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = ptx::lane_id();
  const uint32_t cta_rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int cluster_id = (int)(blockIdx.x / CG), nclusters = (int)(gridDim.x / CG);
  const int ntiles = p.tiles_m * p.tiles_n;

  if (CG == 2) ptx::cluster_sync();    // peer CTA must be resident before any remote barrier traffic

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tensormap(&tmAhi); ptx::prefetch_tensormap(&tmAlo);
    ptx::prefetch_tensormap(&tmBhi); ptx::prefetch_tensormap(&tmBlo);
  }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < Cfg::STAGES; s++) {
      ptx::mbar_init(full_bar(s), CG);      // one arrival per producer CTA (both arrive on the leader's)
      ptx::mbar_init(empty_bar(s), 1);      // one tcgen05.commit (multicast to both CTAs when CG == 2)
    }
    for (int b = 0; b < 2; b++) {
      ptx::mbar_init(tfull_bar(b), 1);          // chain complete (commit, multicast)
      ptx::mbar_init(tempty_bar(b), 8 * CG);    // buffer drained: one arrival per accumulate warp of every CTA (leader's barrier)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<CG>(tmem_slot, Cfg::TMEM_COLS);
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int nkb = p.kblocks;
  const int flush = p.flush_kb;
  const int nchains = (nkb + flush - 1) / flush;    // per tile

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ===================== TMA producer (one elected lane, in every CTA) =====================
      if (ptx::elect_one()) {
        uint32_t it = 0;
        unsigned long long expected = 0;      // cumulative arrivals the grid barrier must have seen
        int round = 0;
        bool lockstep = true;                 // cleared by the first timeout: the CTAs are not all co-resident
        for (int tile = cluster_id; tile < ntiles; tile += nclusters, round++) {
          if (round > 0 && p.sync_counter != nullptr) {
            // grid barrier (bounded spin): every CTA that has a tile in this round arrives once
            const int rem = ntiles - round * nclusters;
            expected += (unsigned long long)CG * (unsigned)(rem < nclusters ? rem : nclusters);
            atomicAdd(p.sync_counter, 1u);
            const long long t0 = clock64();
            while (lockstep && *((volatile unsigned*)p.sync_counter) < (unsigned)expected) {
              // ~0.2 ms: never deadlock if a CTA is not co-resident (e.g. a communication kernel holds some SMs);
              // after ONE timeout this CTA stops waiting for the rest of the launch (it still arrives, so the others'
              // counts stay right) — the schedule degrades to the free-running one instead of paying 0.2 ms per wave
              if (clock64() - t0 > 400000) { lockstep = false; break; }
              __nanosleep(64);
            }
          }
          int tm, tn;
          tile_coords(p, tile, &tm, &tn);
          const int row0 = tm * Cfg::TILE_M + (int)cta_rank * 128;     // first A row staged by this CTA
          const int col0 = tn * Cfg::TILE_N;                           // first B row (output column) of the tile
          for (int kb = 0; kb < nkb; kb++, it++) {
            const int s = it % Cfg::STAGES;
            const uint32_t ph = (it / Cfg::STAGES) & 1u;
            ptx::mbar_wait(empty_bar(s), ph ^ 1u);
            const uint32_t sb = stage_base(s);
            const int kc = kb * BK;
            if constexpr (CG == 1) {
              const uint32_t fb = full_bar(s);
              ptx::mbar_arrive_expect_tx(fb, Cfg::STAGE_BYTES);
              ptx::tma_load_2d(sb + OFF_AHI, &tmAhi, fb, kc, row0);
              ptx::tma_load_2d(sb + OFF_BHI, &tmBhi, fb, kc, col0);
              ptx::tma_load_2d(sb + OFF_BHI + Cfg::BOX_BYTES, &tmBhi, fb, kc, col0 + 128);
              ptx::tma_load_2d(sb + OFF_ALO, &tmAlo, fb, kc, row0);
              ptx::tma_load_2d(sb + OFF_BLO, &tmBlo, fb, kc, col0);
              ptx::tma_load_2d(sb + OFF_BLO + Cfg::BOX_BYTES, &tmBlo, fb, kc, col0 + 128);
            } else {
              // all transaction bytes of both CTAs land on the LEADER's full barrier
              const uint32_t fb = ptx::mapa(full_bar(s), 0);
              if (leader) ptx::mbar_arrive_expect_tx(full_bar(s), 2 * Cfg::STAGE_BYTES);
              else ptx::mbar_arrive_cluster(fb);
              const int b0 = col0 + (int)cta_rank * 128;     // this CTA supplies B rows [rank*128, +128) of the 256 columns
              ptx::tma_load_2d_pair(sb + OFF_AHI, &tmAhi, fb, kc, row0);
              ptx::tma_load_2d_pair(sb + OFF_BHI, &tmBhi, fb, kc, b0);
              ptx::tma_load_2d_pair(sb + OFF_ALO, &tmAlo, fb, kc, row0);
              ptx::tma_load_2d_pair(sb + OFF_BLO, &tmBlo, fb, kc, b0);
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===================== UMMA issuer (leader CTA only, one elected lane) =====================
      if (leader && ptx::elect_one()) {
        const uint64_t dhi = ptx::umma_desc_hi(Cfg::SBO, Cfg::ROW_BYTES);
        const uint32_t idesc = ptx::umma_idesc_tf32(128 * CG, 256);
        uint32_t it = 0, chain = 0;
        for (int tile = cluster_id; tile < ntiles; tile += nclusters) {
          int kb = 0;
          for (int c = 0; c < nchains; c++, chain++) {
            const int buf = chain & 1;
            ptx::mbar_wait(tempty_bar(buf), ((chain >> 1) & 1u) ^ 1u);    // accumulate warps drained this buffer
            ptx::tc_fence_after();
            const uint32_t d = tmem_base + (uint32_t)buf * 256u;
            const int kb_end = (kb + flush < nkb) ? kb + flush : nkb;
            for (bool first = true; kb < kb_end; kb++, it++) {
              const int s = it % Cfg::STAGES;
              const uint32_t ph = (it / Cfg::STAGES) & 1u;
              ptx::mbar_wait(full_bar(s), ph);
              ptx::tc_fence_after();
              const uint32_t sb = stage_base(s);
#pragma unroll
              for (int k8 = 0; k8 < BK / 8; k8++) {
                const uint32_t koff = k8 * 32;                       // 8 tf32 = 32 bytes along K inside the swizzle atom
                const uint64_t a_hi = ptx::umma_desc(dhi, sb + OFF_AHI + koff);
                const uint64_t a_lo = ptx::umma_desc(dhi, sb + OFF_ALO + koff);
                const uint64_t b_hi = ptx::umma_desc(dhi, sb + OFF_BHI + koff);
                const uint64_t b_lo = ptx::umma_desc(dhi, sb + OFF_BLO + koff);
                ptx::umma_tf32<CG>(d, a_lo, b_hi, idesc, first ? 0u : 1u);   // small terms first
                ptx::umma_tf32<CG>(d, a_hi, b_lo, idesc, 1u);
                ptx::umma_tf32<CG>(d, a_hi, b_hi, idesc, 1u);
                first = false;
              }
              ptx::umma_commit<CG>(empty_bar(s));                    // frees the stage in both CTAs once the MMAs retire
            }
            ptx::umma_commit<CG>(tfull_bar(buf));                    // chain complete -> accumulate warps
          }
        }
      }
    }
  } else {
    // ===================== accumulate / epilogue warps 4..11 =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;              // which 128 of the 256 columns
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)half * 128u;
    const uint32_t tempty_leader = (CG == 2) ? ptx::mapa(tempty_bar(0), 0) : tempty_bar(0);
    uint32_t chain = 0;
    // fused all-gather: a finished tile is stored to the LOCAL copy of C in the epilogue; its copies to the peers are
    // trickled out between the chain drains of the NEXT tile (this thread re-reads its own row from L2 and posts it to
    // the 7 peers, a few elements per chain).  Storing all 8 copies in the epilogue made every CTA burst ~1 MB onto
    // NVLink at the same moment (the per-wave grid barrier keeps the CTAs in lockstep) and cost 21 % at 8 GPUs.
    const float* fwd_src = nullptr;               // this thread's row segment of the previous tile (local copy)
    int64_t fwd_off = 0;                          // its offset inside C (same in every copy)
    int fwd_n = 0, fwd_pos = 0;                   // valid columns / columns already forwarded
    auto forward = [&](int count) {
      for (int e = 0; e < count && fwd_pos < fwd_n; e++, fwd_pos++) {
        const int64_t o = fwd_off + (int64_t)fwd_pos * p.csC;
        const float v = fwd_src[(int64_t)fwd_pos * p.csC];
#pragma unroll
        for (int g = 0; g < 8; g++)
          if (g < p.npeers && g != p.self) p.peer[g][o] = v;
      }
    };
    for (int tile = cluster_id; tile < ntiles; tile += nclusters) {
      float acc[128];
#pragma unroll
      for (int i = 0; i < 128; i++) acc[i] = 0.f;
      for (int c = 0; c < nchains; c++, chain++) {
        const int buf = chain & 1;
        if (p.npeers > 1) forward((int)(((int64_t)(c + 1) * 128 + nchains - 1) / nchains) - fwd_pos);   // evenly over the tile
        ptx::mbar_wait(tfull_bar(buf), (chain >> 1) & 1u);
        ptx::tc_fence_after();
#pragma unroll
        for (int j = 0; j < 4; j++) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(tlane + (uint32_t)buf * 256u + (uint32_t)j * 32u, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) acc[j * 32 + i] = __fadd_rn(acc[j * 32 + i], __uint_as_float(r[i]));
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) ptx::mbar_arrive_cluster(tempty_leader + 8u * buf);
          else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(buf)) : "memory");
        }
      }
      // ---- epilogue: alpha/beta, strided stores (lanes run along C's unit-stride dimension); overlaps the next tile
      int tm, tn;
      tile_coords(p, tile, &tm, &tn);
      const int64_t m = (int64_t)tm * Cfg::TILE_M + (int64_t)cta_rank * 128 + q * 32 + (int)lane;
      const int64_t col0 = (int64_t)tn * Cfg::TILE_N;
      if (p.npeers > 0) {
        // GEMM -> all-gather in one kernel: this tile to the local copy now, to the peers during the next tile
        forward(128);                              // whatever is left of the previous tile
        fwd_n = 0; fwd_pos = 0;
        if (m < p.M) {
          const int64_t n0 = col0 + half * 128;
          const int64_t orow = m * p.rsC + n0 * p.csC;
          float* crow = p.peer[p.self] + orow;
          const float alpha = p.alpha;
          int nvalid = 0;
#pragma unroll
          for (int i = 0; i < 128; i++) {           // full unroll: acc[] must stay in registers
            if (n0 + i < p.N) { crow[i * p.csC] = epilogue_value<float>(alpha, acc[i], 0.f, 0.f); nvalid = i + 1; }
          }
          fwd_src = crow; fwd_off = orow; fwd_n = nvalid;
        }
      } else if (m < p.M) {
        float* crow = p.C + m * p.rsC;
        const float alpha = p.alpha, beta = p.beta;
        const float bl = p.bias_lane ? p.bias_lane[m] : 0.f;
#pragma unroll
        for (int i = 0; i < 128; i++) {
          const int64_t n = col0 + half * 128 + i;
          if (n < p.N) {
            float* pc = crow + n * p.csC;
            const float cold = (beta != 0.f) ? *pc : 0.f;
            float v = epilogue_value<float>(alpha, acc[i], beta, cold);
            if (p.bias_lane) v = __fadd_rn(v, bl);
            if (p.bias_col) v = __fadd_rn(v, p.bias_col[n]);
            *pc = v;
          }
        }
      }
    }
    if (p.npeers > 1) forward(128);                // the last tile's copies
  }

  // teardown: nobody may leave (or free TMEM) while the pair still has traffic in flight
  __syncwarp();
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_TmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);
static PFN_TmapEncodeTiled g_encode = nullptr;
static int g_tc_state = 0;   // 0 unknown, 1 available, -1 unavailable

bool gemm_f32_tc_available() {
  if (g_tc_state == 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      g_tc_state = -1; return false;
    }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (major != 10 || cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !fn) {
      g_tc_state = -1; return false;
    }
    g_encode = (PFN_TmapEncodeTiled)fn;
    g_tc_state = 1;
  }
  return g_tc_state == 1;
}

static int make_tmap(CUtensorMap* tm, float* base, int64_t rows, int64_t kpad, int bk) {
  cuuint64_t gdim[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)kpad * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)bk, 128u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return AM_ERR_CUDA; }
  return AM_OK;
}

// Dense (unswizzled) float32 tensor map of rank 2..4 for the staging loads of the conv kernels: dims / box innermost
// first, strides_bytes[i] = byte stride of dimension i+1 (multiples of 16), zero fill outside the tensor.
int make_tmap_f32_nd(CUtensorMap* tm, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
  if (!gemm_f32_tc_available()) { set_last_error("TMA needs a compute-capability 10.x device"); return AM_ERR_UNSUPPORTED; }
  cuuint64_t gdim[5]; cuuint64_t gstr[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; i++) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1u; }
  for (int i = 0; i + 1 < rank; i++) gstr[i] = strides_bytes[i];
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled (rank %d) failed (%d)", rank, (int)r); return AM_ERR_CUDA; }
  return AM_OK;
}

// same, 128-byte swizzle (the box's inner dimension must be 32 floats): lands tiles in the K-major SW128 layout UMMA reads
int make_tmap_f32_nd_sw128(CUtensorMap* tm, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box) {
  if (!gemm_f32_tc_available()) { set_last_error("TMA needs a compute-capability 10.x device"); return AM_ERR_UNSUPPORTED; }
  cuuint64_t gdim[5]; cuuint64_t gstr[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; i++) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1u; }
  for (int i = 0; i + 1 < rank; i++) gstr[i] = strides_bytes[i];
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled (rank %d, sw128) failed (%d)", rank, (int)r); return AM_ERR_CUDA; }
  return AM_OK;
}

// Dense 2-D float64 tensor map (inner dimension first) for the TMA-fed DMMA kernel; zero fill outside the tensor.
int make_tmap_f64_2d(CUtensorMap* tm, const double* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                     uint32_t box_inner, uint32_t box_outer) {
  if (!gemm_f32_tc_available()) { set_last_error("TMA needs a compute-capability 10.x device"); return AM_ERR_UNSUPPORTED; }
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {outer_stride_bytes};
  cuuint32_t bx[2] = {box_inner, box_outer};
  cuuint32_t es[2] = {1u, 1u};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled (f64) failed (%d)", (int)r); return AM_ERR_CUDA; }
  return AM_OK;
}

template <int CG>
static int launch_tc(cudaStream_t st, const CUtensorMap* tms, const TcArgs& args) {
  using Cfg = TcCfg<CG>;
  auto kern = gemm_tf32x3_kernel<CG>;
  AM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  cudaLaunchConfig_t cfg{};
  const int ntiles = args.tiles_m * args.tiles_n;
  const int resident = sm_count() / CG;                          // one CTA (pair) per SM (pair)
  const int nclusters = ntiles < resident ? ntiles : resident;
  cfg.gridDim = dim3((unsigned)(nclusters * CG), 1, 1);
  cfg.blockDim = dim3(Cfg::NTHREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  AM_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tms[0], tms[1], tms[2], tms[3], args));
  g_launch_count++;
  return AM_OK;
}

// A packed operand: two K-major planes (hi, lo) of [Rpad][Kpad] floats, tf32-rounded.
struct PackedF32 {
  float* hi; float* lo;
  int64_t R, K, Rpad, Kpad;
  bool owned;          // planes were cudaMalloc'd for this object (am_pack_f32) vs. workspace
};

static int64_t pad_rows(int64_t r) { return round_up(r, 512); }   // either GEMM role (256-row or 512-row tiles)
static int64_t pad_k(int64_t k) { return round_up(k, 32); }

static int pack_into(cudaStream_t st, const float* X, int64_t R, int64_t K, int64_t r_stride, int64_t k_stride,
                     PackedF32* out) {
  if (out->Rpad / 32 > 65535) { set_last_error("gemm_f32_tc: dimension too large for the pack grid"); return AM_ERR_INVALID; }
  const bool kf = iabs64(k_stride) <= iabs64(r_stride);
  const bool vec_ok = (reinterpret_cast<uintptr_t>(X) & 15) == 0 && R % 4 == 0 && K % 4 == 0 && out->Rpad % 64 == 0 &&
                      (kf ? (k_stride == 1 && r_stride % 4 == 0) : (r_stride == 1 && k_stride % 4 == 0));
  if (vec_ok && !tuning(kTunePackScalar)) {
    // rows r >= R (padding) are produced as zeros by the bounds test; k in [K, Kpad) likewise (Kpad - K < 32, K % 4 == 0)
    split_pack_vec_kernel<<<dim3((unsigned)ceil_div(out->Kpad, 64), (unsigned)(out->Rpad / 64)), 256, 0, st>>>(
        X, R, K, r_stride, k_stride, out->hi, out->lo, out->Kpad, kf ? 1 : 0);
    g_launch_count++;
    AM_CUDA_TRY(cudaGetLastError());
    return AM_OK;
  }
  split_pack_kernel<<<dim3((unsigned)(out->Kpad / 32), (unsigned)(out->Rpad / 32)), 256, 0, st>>>(
      X, R, K, r_stride, k_stride, out->hi, out->lo, out->Kpad, iabs64(k_stride) <= iabs64(r_stride));
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  return AM_OK;
}

// P rows -> TMEM lanes (C's unit-stride dimension), Q rows -> TMEM columns.
static int run_packed(cudaStream_t st, int cta_group, const PackedF32& P, const PackedF32& Q, float alpha,
                      float beta, float* C, int64_t strideP, int64_t strideQ, int npeers = 0, float* const* peers = nullptr,
                      int self = 0, const float* bias_p = nullptr, const float* bias_q = nullptr) {
  if (P.Kpad != Q.Kpad || P.K != Q.K) { set_last_error("gemm_f32_tc: packed operands disagree on K"); return AM_ERR_INVALID; }
  const int flush_env = tuning(kTuneTcFlushKb) > 0 ? tuning(kTuneTcFlushKb) : 2;
  const int bk = 32;
  CUtensorMap tms[4];
  int rc;
  if ((rc = make_tmap(&tms[0], P.hi, P.Rpad, P.Kpad, bk)) || (rc = make_tmap(&tms[1], P.lo, P.Rpad, P.Kpad, bk)) ||
      (rc = make_tmap(&tms[2], Q.hi, Q.Rpad, Q.Kpad, bk)) || (rc = make_tmap(&tms[3], Q.lo, Q.Rpad, Q.Kpad, bk)))
    return rc;
  TcArgs args;
  const int group_env = tuning(kTuneTcGroup) != 0 ? tuning(kTuneTcGroup) : 8;   // > 0: groups of M-tiles (M fastest), < 0: groups of N-tiles
  args.M = P.R; args.N = Q.R; args.kblocks = (int)(P.Kpad / bk); args.flush_kb = flush_env; args.group = group_env;
  args.C = C; args.rsC = strideP; args.csC = strideQ; args.alpha = alpha; args.beta = beta;
  args.npeers = npeers; args.self = self; args.bias_lane = bias_p; args.bias_col = bias_q;
  for (int g = 0; g < 8; g++) args.peer[g] = (g < npeers) ? peers[g] : nullptr;
  // grid-barrier counter of the persistent schedule: one slot of a small ring, zeroed in stream order
  args.sync_counter = nullptr;
  if (tuning(kTuneTcSync)) {
    static std::atomic<unsigned> ring{0};
    void* base = nullptr;
    if ((rc = workspace(kWsMisc, 64 * sizeof(int) + 1024, &base))) return rc;
    unsigned* ctr = (unsigned*)base + 64 + (ring++ % 64);
    AM_CUDA_TRY(cudaMemsetAsync(ctr, 0, sizeof(unsigned), st));
    args.sync_counter = ctr;
  }
  args.tiles_n = (int)ceil_div(Q.R, 256);
  if (cta_group == 2) {
    args.tiles_m = (int)ceil_div(P.R, 256);
    return launch_tc<2>(st, tms, args);
  }
  args.tiles_m = (int)ceil_div(P.R, 128);
  return launch_tc<1>(st, tms, args);
}

// prepackedB != nullptr: B was split/packed by the caller (a handle from pack_f32 / pack_f32_view for B[K,N] in the
// "b" role, i.e. rows = n) and is reused as is — only A is packed here (row chunks of one host-buffer product).
int gemm_f32_tc(cudaStream_t st, int cta_group, int64_t M, int64_t N, int64_t K, float alpha, const float* A,
                int64_t rsA, int64_t csA, const float* B, int64_t rsB, int64_t csB, float beta, float* C,
                int64_t rsC, int64_t csC, const void* prepackedB, const float* bias_col) {
  if (!gemm_f32_tc_available()) { set_last_error("tcgen05 path needs a compute-capability 10.x device"); return AM_ERR_UNSUPPORTED; }
  if (M >= (1ll << 31) - 1024 || N >= (1ll << 31) - 1024 || K >= (1ll << 31) - 64) { set_last_error("gemm_f32_tc: dimension too large"); return AM_ERR_INVALID; }
  // Operands in (panel-row stride, k stride) form: A panel rows = m, B panel rows = n.
  PackedF32 pa{nullptr, nullptr, M, K, pad_rows(M), pad_k(K), false};
  PackedF32 pb{nullptr, nullptr, N, K, pad_rows(N), pad_k(K), false};
  void *wsA = nullptr, *wsB = nullptr;
  int rc = workspace(kWsSplitA, (size_t)(2 * pa.Rpad * pa.Kpad) * sizeof(float), &wsA);
  if (rc) return rc;
  pa.hi = (float*)wsA; pa.lo = pa.hi + pa.Rpad * pa.Kpad;
  if ((rc = pack_into(st, A, M, K, rsA, csA, &pa))) return rc;
  if (prepackedB) {
    pb = *(const PackedF32*)prepackedB;
    if (pb.R != N || pb.K != K) { set_last_error("gemm_f32_tc: pre-packed B does not match N, K"); return AM_ERR_INVALID; }
  } else {
    rc = workspace(kWsSplitB, (size_t)(2 * pb.Rpad * pb.Kpad) * sizeof(float), &wsB);
    if (rc) return rc;
    pb.hi = (float*)wsB; pb.lo = pb.hi + pb.Rpad * pb.Kpad;
    if ((rc = pack_into(st, B, N, K, csB, rsB, &pb))) return rc;
  }
  // The epilogue's lanes run along the P rows (TMEM lanes): make that C's unit-stride dimension.
  // Column-major C (rs == 1, the CudaTensor default): P = A.  Row-major C: C^T = B^T A^T, P = B.
  // (bias is per column n of C: the Q index when P = A, the lane index when P = B)
  if (iabs64(rsC) <= iabs64(csC)) return run_packed(st, cta_group, pa, pb, alpha, beta, C, rsC, csC, 0, nullptr, 0, nullptr, bias_col);
  return run_packed(st, cta_group, pb, pa, alpha, beta, C, csC, rsC, 0, nullptr, 0, bias_col, nullptr);
}

// ---- pre-packed operands (laser's gemm_prepacked.nim:276-293 on the device): pack once, multiply many
int pack_f32(cudaStream_t st, int64_t R, int64_t K, const float* X, int64_t r_stride, int64_t k_stride, void** handle) {
  if (!gemm_f32_tc_available()) { set_last_error("tcgen05 path needs a compute-capability 10.x device"); return AM_ERR_UNSUPPORTED; }
  if (R <= 0 || K <= 0 || !X || !handle) { set_last_error("am_pack_f32: bad argument"); return AM_ERR_INVALID; }
  PackedF32* p = new PackedF32{nullptr, nullptr, R, K, pad_rows(R), pad_k(K), true};
  cudaError_t e = cudaMalloc((void**)&p->hi, (size_t)(2 * p->Rpad * p->Kpad) * sizeof(float));
  if (e != cudaSuccess) { delete p; return cuda_fail(e, "cudaMalloc(packed operand)"); }
  p->lo = p->hi + p->Rpad * p->Kpad;
  int rc = pack_into(st, X, R, K, r_stride, k_stride, p);
  if (rc) { cudaFree(p->hi); delete p; return rc; }
  *handle = p;
  return AM_OK;
}
// packed operand over caller-provided device memory (stream-ordered allocations of the host-buffer GEMM): `planes`
// holds packed_floats_f32(R, K) floats; the handle does not own them (packed_free_f32 only drops the descriptor)
int64_t packed_floats_f32(int64_t R, int64_t K) { return 2 * pad_rows(R) * pad_k(K); }
int pack_f32_view(cudaStream_t st, int64_t R, int64_t K, const float* X, int64_t r_stride, int64_t k_stride, float* planes,
                  void** handle) {
  if (!gemm_f32_tc_available()) { set_last_error("tcgen05 path needs a compute-capability 10.x device"); return AM_ERR_UNSUPPORTED; }
  if (R <= 0 || K <= 0 || !X || !handle || !planes) { set_last_error("pack_f32_view: bad argument"); return AM_ERR_INVALID; }
  PackedF32* p = new PackedF32{planes, planes + pad_rows(R) * pad_k(K), R, K, pad_rows(R), pad_k(K), false};
  int rc = pack_into(st, X, R, K, r_stride, k_stride, p);
  if (rc) { delete p; return rc; }
  *handle = p;
  return AM_OK;
}
// handle over planes that already hold a packed operand (e.g. a K slice of B received from a peer GPU)
int packed_wrap_f32(int64_t R, int64_t K, float* planes, void** handle) {
  if (R <= 0 || K <= 0 || !handle || !planes) { set_last_error("packed_wrap_f32: bad argument"); return AM_ERR_INVALID; }
  *handle = new PackedF32{planes, planes + pad_rows(R) * pad_k(K), R, K, pad_rows(R), pad_k(K), false};
  return AM_OK;
}
int repack_f32(cudaStream_t st, void* handle, const float* X, int64_t r_stride, int64_t k_stride) {
  PackedF32* p = (PackedF32*)handle;
  if (!p || !X) { set_last_error("am_repack_f32: bad argument"); return AM_ERR_INVALID; }
  return pack_into(st, X, p->R, p->K, r_stride, k_stride, p);
}
int packed_free_f32(void* handle) {
  PackedF32* p = (PackedF32*)handle;
  if (!p) return AM_OK;
  if (p->owned && p->hi) cudaFree(p->hi);
  delete p;
  return AM_OK;
}
int gemm_packed_f32(cudaStream_t st, float alpha, const void* hA, const void* hB, float beta, float* C, int64_t rsC,
                    int64_t csC) {
  const PackedF32* a = (const PackedF32*)hA; const PackedF32* b = (const PackedF32*)hB;
  if (!a || !b || !C) { set_last_error("am_gemm_packed_f32: bad argument"); return AM_ERR_INVALID; }
  if (iabs64(rsC) <= iabs64(csC)) return run_packed(st, 2, *a, *b, alpha, beta, C, rsC, csC);
  return run_packed(st, 2, *b, *a, alpha, beta, C, csC, rsC);
}

// C <- alpha*A*B written to EVERY peer's copy of C (peers[g] = address of C's element (0,0) in GPU g's buffer as
// mapped into this process, own copy included): the row-sharded GEMM and the all-gather of its result in one kernel.
int gemm_packed_f32_bcast(cudaStream_t st, float alpha, const void* hA, const void* hB, int npeers, float* const* peers,
                          int self, int64_t rsC, int64_t csC) {
  const PackedF32* a = (const PackedF32*)hA; const PackedF32* b = (const PackedF32*)hB;
  if (!a || !b || !peers || npeers < 1 || npeers > 8 || self < 0 || self >= npeers) { set_last_error("am_gemm_packed_f32_bcast: bad argument (1..8 peers)"); return AM_ERR_INVALID; }
  for (int g = 0; g < npeers; g++) if (!peers[g]) { set_last_error("am_gemm_packed_f32_bcast: null peer pointer"); return AM_ERR_INVALID; }
  if (iabs64(rsC) <= iabs64(csC)) return run_packed(st, 2, *a, *b, alpha, 0.f, peers[self], rsC, csC, npeers, peers, self);
  return run_packed(st, 2, *b, *a, alpha, 0.f, peers[self], csC, rsC, npeers, peers, self);
}

}  // namespace am
