// float32 conv2d data gradient on the 5th-gen tensor cores in the reference's own "GEMM, then col2im" form
// (conv.nim:131-139: gcol = Wcol^T * gout[i]; gin[i] = col2im(gcol)), any stride / padding / dilation:
//
//   D[q, r] = sum_co gout[n, co, q] * W[co, ci, kh, kw]      q = output pixel (M, 128 per tile = TMEM lanes),
//                                                             r = (ci, kh, kw) (N, up to 128 rows per chunk),
//   gin[n, ci, ho*s - p + kh*d, wo*s - p + kw*d] += D[q, r]   (col2im)
//
// The contraction length is only Cout, so the whole job is ONE short accumulation chain per tile (<= 64 products:
// no truncation build-up in the tensor core's fp32 accumulate, see gemm_f32_tc.cu) and the expensive 25x im2col
// expansion of the gather form disappears: the A operand is just gout itself.  3xTF32 split as everywhere else.
//
// Persistent, warp-specialised CTA (768 threads).  A CTA owns one CHUNK of input channels (cpc channels = up to 128
// rows r; its W^T slice stays resident in shared memory) and walks over groups of whole images:
//   warp 0      one TMA load of the CTA's W^T chunk (hi / lo planes, K-major, 128B swizzle)
//   warp 1      UMMA issuer: A from tensor memory, B from shared memory, (Kpad/8)*3 tcgen05.mma per tile
//   warp 2      TMEM allocation (2 accumulators of 128 columns + 2 operand stages of 2*Kpad columns)
//   warps 4-7   operand warps: thread q loads gout[n, :, q] (lanes = consecutive pixels: coalesced), splits into
//               tf32 hi / lo and writes its TMEM lane with tcgen05.st
//   warps 8-23  col2im warps: tcgen05.ld the tile's D row, park it in shared memory (col_s[r][q]; the accumulator is
//               free again at this point), then every thread gathers the taps that land on ITS input pixels and
//               accumulates them in a shared-memory image (fixed order: deterministic); whole images are flushed
//               to HBM with coalesced stores
#include <cuda.h>
#include <cstdio>
#include <cstdlib>

#include "am_common.cuh"
#include "gemm_dispatch.h"
#include "ptx_sm100.cuh"

namespace am {

// Per-role wait-cycle counters (tuning knob "convtc_debug") cost two clock reads around every mbarrier wait — in the
// single-thread MMA issue loop that is ~150 cycles per k block.  They are compiled out unless AM_CONVTC_PROFILE is set:
// pclk() is then a constant and the counters fold away.
#ifndef AM_CONVTC_PROFILE
#define AM_CONVTC_PROFILE 0
#endif
#ifndef AM_PCLK_DEFINED
#define AM_PCLK_DEFINED
__device__ __forceinline__ long long pclk() {
#if AM_CONVTC_PROFILE
  return clock64();
#else
  return 0;
#endif
}
#endif


// grad_output as pre-split tf32 hi / lo planes ([N][CO][HO*WO] each, lo right behind hi) for the TMA-fed weight-gradient
// kernel: can they be used for this call?  (rows of 16-byte multiples, aligned base, at most 1 GiB of planes, knob on)
static bool gout_planes_ok(const am_conv2d_desc& d, int64_t HWo, const float* grad_output, int64_t* total) {
  *total = d.N * d.Cout * HWo;
  return tuning(kTuneConvTcWgradTma) != 0 && HWo % 4 == 0 && (reinterpret_cast<uintptr_t>(grad_output) & 15) == 0 &&
         d.N < (1ll << 31) && *total > 0 && *total <= (1ll << 27);
}

struct DgradTcArgs {
  const float* gout;    // [N][CO][HO][WO]
  float* gin;           // [N][C][H][W]
  int64_t N;
  int C, H, W, CO, HO, WO, kH, kW, padH, padW, sH, sW, dH, dW;
  int Kpad;             // Cout padded to 32 or 64 (k blocks of 32)
  int cpc;              // input channels per chunk (cpc * kH * kW <= 128)
  int nchunks;          // ceil(C / cpc)
  int ipg;              // whole images per group (ipg * HO * WO <= 128, or 1 when an image needs several tiles)
  int tpg;              // 128-pixel tiles per group
  int64_t ngroups;      // ceil(N / ipg)
  int padded;           // 1: single-tile groups, unit stride/dilation: col_s rows are zero-padded images (no tap predicates)
  int WP, colL;         // padded row pitch WO + kW - 1, floats per col_s column
  long long* dbg;       // optional per-role cycle counters of CTA 0 (AM_CONVTC_DEBUG=1)
};

__global__ void dgrad_tc_pack_weights_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo,
                                             int C, int CO, int KK, int cpc, int nchunks, int Kpad) {
  // row (chunk, r): ci = chunk*cpc + r / KK, tap = r % KK (zero rows beyond cpc*KK or C); column co (zero beyond CO)
  const int64_t total = (int64_t)nchunks * 128 * Kpad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Kpad);
    const int row = (int)(i / Kpad), chunk = row >> 7, r = row & 127;
    const int cl = r / KK, tap = r - cl * KK, ci = chunk * cpc + cl;
    float v = 0.f;
    if (cl < cpc && ci < C && co < CO) v = w[((int64_t)co * C + ci) * KK + tap];
    float h = v, l = 0.f;
    if (isfinite(v)) { h = ptx::to_tf32_rna(v); l = ptx::to_tf32_rna(v - h); if (!isfinite(h)) { h = v; l = 0.f; } }
    hi[i] = h;
    lo[i] = l;
  }
}

// column length of the zero-padded col_s layout: a compile-time constant so that tap kw of a kernel row is an
// immediate offset (kw * 1 KiB) from the row's address
constexpr int kPadColL = 257;

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(768, 1)
conv_dgrad_tc_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo, const DgradTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nkb = a.Kpad >> 5;
  const int KK = a.kH * a.kW;
  const int HWi = a.H * a.W, HWo = a.HO * a.WO;
  const uint32_t plane_bytes = (uint32_t)nkb * 16384u;                   // 128 rows x Kpad floats, k blocks of 32
  const uint32_t OFF_BHI = 0, OFF_BLO = plane_bytes;
  const uint32_t col_base = smem_base + 2u * plane_bytes;                 // col_s[cpc*KK <= 128][colL]
  const uint32_t col_bytes = (uint32_t)(a.cpc * KK) * (uint32_t)a.colL * 4u;
  const uint32_t img_base = col_base + col_bytes;                         // gin_s[cpc][ipg][H*W]
  const uint32_t img_floats = (uint32_t)(a.cpc * a.ipg * HWi);
  const uint32_t bar_base = (img_base + (img_floats + 4u) * 4u + 15u) & ~15u;            // + the zero word
  const uint32_t b_full = bar_base;
  auto a_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto a_empty = [&](int s) { return bar_base + 8u * (3 + s); };
  auto d_full = [&](int b) { return bar_base + 8u * (5 + b); };
  auto d_empty = [&](int b) { return bar_base + 8u * (7 + b); };
  const uint32_t tmem_slot = bar_base + 8u * 9;

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = ptx::lane_id();
  const int chunk = (int)blockIdx.x % a.nchunks;
  const int64_t g0 = (int64_t)blockIdx.x / a.nchunks, gstep = (int64_t)gridDim.x / a.nchunks;
  const int ci0 = chunk * a.cpc;
  const int cpc_here = (a.C - ci0 < a.cpc) ? a.C - ci0 : a.cpc;

  if (warp == 0 && ptx::elect_one()) { ptx::prefetch_tensormap(&tmWhi); ptx::prefetch_tensormap(&tmWlo); }
  if (warp == 1 && ptx::elect_one()) {
    ptx::mbar_init(b_full, 1);
    for (int s = 0; s < 2; s++) {
      ptx::mbar_init(a_full(s), 128); ptx::mbar_init(a_empty(s), 1);
      ptx::mbar_init(d_full(s), 1); ptx::mbar_init(d_empty(s), 16);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<1>(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tmem_a0 = tmem_base + 256u;                              // operand stages after the two accumulators

  // register budget (24 warps at 80): control warps 56, operand warps 120, col2im warps 72
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ===================== one-off TMA load of this CTA's W^T chunk =====================
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(b_full, 2 * plane_bytes);
      for (int kb = 0; kb < nkb; kb++) {
        ptx::tma_load_2d(smem_base + OFF_BHI + (uint32_t)kb * 16384u, &tmWhi, b_full, kb * 32, chunk * 128);
        ptx::tma_load_2d(smem_base + OFF_BLO + (uint32_t)kb * 16384u, &tmWlo, b_full, kb * 32, chunk * 128);
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    if (ptx::elect_one()) {
      const uint64_t dhi = ptx::umma_desc_hi(1024, 128);
      const uint32_t idesc = ptx::umma_idesc_tf32(128, 128);
      ptx::mbar_wait(b_full, 0);
      uint32_t it = 0;
      long long w_de = 0, w_af = 0;
      const long long tstart = pclk();
      for (int64_t g = g0; g < a.ngroups; g += gstep) {
        for (int t = 0; t < a.tpg; t++, it++) {
          const int s = it & 1;
          const uint32_t ph = (it >> 1) & 1u;
          { const long long t0_ = pclk(); ptx::mbar_wait(d_empty(s), ph ^ 1u); w_de += pclk() - t0_; }
          { const long long t0_ = pclk(); ptx::mbar_wait(a_full(s), ph); w_af += pclk() - t0_; }
          ptx::tc_fence_after();
          const uint32_t d = tmem_base + 128u * (uint32_t)s;
          const uint32_t a_hi0 = tmem_a0 + (uint32_t)s * 2u * (uint32_t)a.Kpad, a_lo0 = a_hi0 + (uint32_t)a.Kpad;
          for (int k8 = 0; k8 < a.Kpad / 8; k8++) {
            const uint32_t boff = (uint32_t)(k8 >> 2) * 16384u + (uint32_t)(k8 & 3) * 32u;
            const uint64_t b_hi = ptx::umma_desc(dhi, smem_base + OFF_BHI + boff), b_lo = ptx::umma_desc(dhi, smem_base + OFF_BLO + boff);
            ptx::umma_tf32_ts(d, a_lo0 + 8u * k8, b_hi, idesc, k8 ? 1u : 0u);
            ptx::umma_tf32_ts(d, a_hi0 + 8u * k8, b_lo, idesc, 1u);
            ptx::umma_tf32_ts(d, a_hi0 + 8u * k8, b_hi, idesc, 1u);
          }
          ptx::umma_commit<1>(a_empty(s));
          ptx::umma_commit<1>(d_full(s));
        }
      }
      if (a.dbg && blockIdx.x == 0) { a.dbg[0] = pclk() - tstart; a.dbg[1] = w_de; a.dbg[2] = w_af; }
    }
  }
  } else if (warp < 8) {
    // ===================== operand warps: gout pixels -> tf32 hi/lo -> tensor memory =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
    const int q = (int)threadIdx.x - 128;                       // TMEM lane
    const uint32_t t_lane = tmem_a0 + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t it = 0;
    long long w_ae = 0;
    const long long tstart = pclk();
    for (int64_t g = g0; g < a.ngroups; g += gstep) {
      const int64_t n0 = g * a.ipg;
      const int imgs = (a.N - n0 < a.ipg) ? (int)(a.N - n0) : a.ipg;
      const int npix = imgs * HWo;
      for (int t = 0; t < a.tpg; t++, it++) {
        const int s = it & 1;
        const int gq = t * 128 + q;                             // pixel index inside the group
        const bool ok = gq < npix;
        const int il = ok ? gq / HWo : 0, pix = ok ? gq - il * HWo : 0;
        const float* src = a.gout + ((n0 + il) * a.CO) * (int64_t)HWo + pix;
        // all of the pixel's channels in flight at once (one L2 round trip per tile instead of one per 16 channels),
        // issued before waiting for the operand stage so the latency overlaps the previous tile's MMAs
        float v[64];
#pragma unroll
        for (int j = 0; j < 64; j++) v[j] = (ok && j < a.CO) ? __ldg(src + (int64_t)j * HWo) : 0.f;
        { const long long t0_ = pclk(); ptx::mbar_wait(a_empty(s), ((it >> 1) & 1u) ^ 1u); w_ae += pclk() - t0_; }
        ptx::tc_fence_after();
        const uint32_t ta = t_lane + (uint32_t)s * 2u * (uint32_t)a.Kpad;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
          if (c0 < a.Kpad) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; j++) ptx::split_tf32(v[c0 + j], hi[j], lo[j]);
            ptx::tmem_st_32x16(ta + (uint32_t)c0, hi);
            ptx::tmem_st_32x16(ta + (uint32_t)a.Kpad + (uint32_t)c0, lo);

          }
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a_full(s)) : "memory");
      }
    }
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 128) { a.dbg[3] = pclk() - tstart; a.dbg[4] = w_ae; }
  } else {
    // ===================== col2im warps (16 warps = 512 threads) =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int tid = (int)threadIdx.x - 256;                     // 0..511
    const int wq = warp & 3;                                    // TMEM lane quarter this warp may read
    const int cgrp = (warp - 8) >> 2;                           // which 32 of the 128 D columns this warp parks
    const int prow = wq * 32 + (int)lane;                       // pixel (TMEM lane) of this thread in phase 1
    const bool unit = a.sH == 1 && a.sW == 1 && a.dH == 1 && a.dW == 1;
    const int ncols = cpc_here * KK;                            // D columns in use
    const float* col_p = reinterpret_cast<const float*>(smem_raw + (col_base - ptx::smem_u32(smem_raw)));
    float* img_p = reinterpret_cast<float*>(smem_raw + (img_base - ptx::smem_u32(smem_raw)));
    for (uint32_t i = tid; i < img_floats + 4; i += 512) img_p[i] = 0.f;
    int ppos = prow;                                            // where this thread's pixel goes inside a col_s column
    if (a.padded) {
      // zero the padding once: phase 1 only ever writes real pixel positions
      float* cz = const_cast<float*>(col_p);
      for (uint32_t i = tid; i < col_bytes / 4u; i += 512) cz[i] = 0.f;
      const int il = prow / HWo, r = prow - il * HWo, ho = r / a.WO, wo = r - ho * a.WO;
      ppos = (prow < a.ipg * HWo) ? (a.kW - 1) + (il * a.HO + ho) * a.WP + wo : -1;
    }
    // padded mode: the (input pixel, channel) items of a thread are the same for every group -> decode them once
    constexpr int kMaxItems = 4;
    const int nitems_full = a.ipg * HWi * cpc_here;
    const bool pre = a.padded && nitems_full <= kMaxItems * 512;
    int it_src[kMaxItems], it_g[kMaxItems], it_k[kMaxItems];
#pragma unroll
    for (int j = 0; j < kMaxItems; j++) {
      it_src[j] = 0; it_g[j] = 0; it_k[j] = -1;
      const int item = tid + 512 * j;
      if (pre && item < nitems_full) {
        const int per_c = a.ipg * HWi;
        const int cl = item / per_c, o = item - cl * per_c;
        const int il = o / HWi, rem = o - il * HWi, h = rem / a.W, w = rem - h * a.W;
        int kh_lo = h + a.padH - (a.HO - 1); if (kh_lo < 0) kh_lo = 0;
        int kh_hi = h + a.padH; if (kh_hi > a.kH - 1) kh_hi = a.kH - 1;
        it_src[j] = cl * KK * a.colL + (a.kW - 1) + (il * a.HO + h + a.padH) * a.WP + (w + a.padW);
        it_g[j] = (il * a.C + ci0 + cl) * HWi + rem;
        it_k[j] = kh_lo | (kh_hi << 8) | (il << 16);
      }
    }
    uint32_t it = 0;
    long long w_df = 0, w_p1 = 0, w_p2 = 0, w_fl = 0;
    const long long tstart = pclk();
    for (int64_t g = g0; g < a.ngroups; g += gstep) {
      const int64_t n0 = g * a.ipg;
      const int imgs = (a.N - n0 < a.ipg) ? (int)(a.N - n0) : a.ipg;
      const int npix = imgs * HWo;
      const int nout = imgs * HWi;                              // input pixels per channel in this group
      for (int t = 0; t < a.tpg; t++, it++) {
        const int s = it & 1;
        const int q_lo = t * 128, q_hi = (q_lo + 128 < npix) ? q_lo + 128 : npix;      // group pixels this tile holds
        const bool whole = (q_lo == 0 && q_hi == npix);         // the tile holds every pixel of the group: no range test
        { const long long t0_ = pclk(); ptx::mbar_wait(d_full(s), (it >> 1) & 1u); w_df += pclk() - t0_; }
        const long long tp1 = pclk();
        ptx::tc_fence_after();
        const uint32_t td = tmem_base + ((uint32_t)(wq * 32) << 16) + 128u * (uint32_t)s;
        // phase 1: D -> col_s[column][pixel] (lanes = consecutive pixels: conflict-free)
        if (cgrp * 32 < ncols) {
          uint32_t r32[32];                                      // this warp's 32 columns in one go (reads past ncols are harmless)
          ptx::tmem_ld_32x32(td + (uint32_t)(cgrp * 32), r32);
          ptx::tmem_ld_wait();
          if (ppos >= 0) {
            const int nleft = ncols - cgrp * 32;                 // columns of this group that exist
            if (a.padded) {                                      // constant column pitch: one address + immediate offsets
              uint32_t* dst = reinterpret_cast<uint32_t*>(const_cast<float*>(col_p)) + cgrp * 32 * kPadColL + ppos;
#pragma unroll
              for (int e = 0; e < 32; e++)
                if (e < nleft) dst[e * kPadColL] = r32[e];
            } else {
#pragma unroll
              for (int e = 0; e < 32; e++)
                if (e < nleft) asm volatile("st.shared.b32 [%0], %1;" ::"r"(col_base + (uint32_t)((cgrp * 32 + e) * a.colL + ppos) * 4u), "r"(r32[e]) : "memory");
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(d_empty(s)) : "memory");   // accumulator free again
        bar_sync_named(1, 512);
        const long long tp2 = pclk();
        w_p1 += tp2 - tp1;
        // phase 2: the input pixels this tile can touch (a row range when an image spans several tiles) x the chunk's
        // channels are dealt out to the 512 threads; each sums the taps that land on its pixel in a fixed order
        // (kh, kw ascending; tiles ascending) -> deterministic
        int o_lo = 0, o_hi = nout;
        if (a.ipg == 1) {
          const int ho_lo = q_lo / a.WO, ho_hi = (q_hi - 1) / a.WO;
          int h_lo = ho_lo * a.sH - a.padH, h_hi = ho_hi * a.sH - a.padH + (a.kH - 1) * a.dH;
          if (h_lo < 0) h_lo = 0;
          if (h_hi > a.H - 1) h_hi = a.H - 1;
          o_lo = h_lo * a.W; o_hi = (h_hi + 1) * a.W;
          if (o_hi < o_lo) o_hi = o_lo;
        }
        const int span = q_hi - q_lo;
        const int cstride = KK * a.colL;                         // floats between the same tap of consecutive channels
        if (pre) {
          // decoded items: rows of kW unconditional loads from the zero-padded column, result straight to HBM
          float* gdst = a.gin + n0 * a.C * (int64_t)HWi;
#pragma unroll
          for (int j = 0; j < kMaxItems; j++) {
            const int k = it_k[j];
            if (k >= 0 && (k >> 16) < imgs) {
              const float* src0 = col_p + it_src[j];
              float sum = 0.f;
              const int rstride = a.kW * kPadColL - a.WP;
              for (int kh = k & 0xff; kh <= ((k >> 8) & 0xff); kh++) {
                const float* src = src0 + kh * rstride;
                float x[8];
#pragma unroll
                for (int kw = 0; kw < 8; kw++) x[kw] = (kw < a.kW) ? src[kw * (kPadColL - 1)] : 0.f;   // immediate offsets
#pragma unroll
                for (int kw = 0; kw < 8; kw++) sum = __fadd_rn(sum, x[kw]);
              }
              gdst[it_g[j]] = sum;
            }
          }
        } else if (a.padded) {
          // (input pixel, channel) items over all 512 threads.  Inside a zero-padded column the source pixel of tap
          // (kh, kw) sits at base - kh*WP - kw, and horizontal misses read padding: no per-tap tests at all.
          const int nitems = nout * cpc_here;
          for (int item = tid; item < nitems; item += 512) {
            const int cl = item / nout, o = item - cl * nout;
            const int il = o / HWi, rem = o - il * HWi, h = rem / a.W, w = rem - h * a.W;
            int kh_lo = h + a.padH - (a.HO - 1); if (kh_lo < 0) kh_lo = 0;
            int kh_hi = h + a.padH; if (kh_hi > a.kH - 1) kh_hi = a.kH - 1;
            const float* src0 = col_p + cl * cstride + (a.kW - 1) + (il * a.HO + h + a.padH) * a.WP + (w + a.padW);
            float sum = 0.f;
            for (int kh = kh_lo; kh <= kh_hi; kh++) {
              const float* src = src0 + kh * (a.kW * a.colL - a.WP);
              float x[8];
#pragma unroll
              for (int kw = 0; kw < 8; kw++) x[kw] = (kw < a.kW) ? src[kw * (a.colL - 1)] : 0.f;
#pragma unroll
              for (int kw = 0; kw < 8; kw++) sum = __fadd_rn(sum, x[kw]);
            }
            float* ga = img_p + cl * a.ipg * HWi + o;
            *ga = __fadd_rn(*ga, sum);
          }
        } else
        for (int o = o_lo + tid; o < o_hi; o += 512) {
          const int il = o / HWi, rem = o - il * HWi, h = rem / a.W, w = rem - h * a.W;
          int kh_lo = 0, kh_hi = a.kH - 1, kw_lo = 0, kw_hi = a.kW - 1;
          const int base0 = il * HWo + (h + a.padH) * a.WO + (w + a.padW) - q_lo;   // unit case: tile pixel of tap (0,0)
          if (unit) {                                            // taps whose source pixel exists: a contiguous box
            kh_lo = h + a.padH - (a.HO - 1); if (kh_lo < 0) kh_lo = 0;
            if (h + a.padH < kh_hi) kh_hi = h + a.padH;
            kw_lo = w + a.padW - (a.WO - 1); if (kw_lo < 0) kw_lo = 0;
            if (w + a.padW < kw_hi) kw_hi = w + a.padW;
          }
          // taps outer, up to 8 channels inner: one address per tap, 8 independent loads + adds (sum order per
          // element: kh, kw ascending)
          for (int c0 = 0; c0 < cpc_here; c0 += 8) {
            float sum[8];
#pragma unroll
            for (int c = 0; c < 8; c++) sum[c] = 0.f;
            const float* cb = col_p + c0 * cstride;
            const int nc = cpc_here - c0;
            for (int kh = kh_lo; kh <= kh_hi; kh++) {
              int ho_off;                                        // ho * WO of the source pixel, or < 0 when there is none
              if (unit) {
                ho_off = 0;
              } else {
                const int hh = h + a.padH - kh * a.dH;
                if (hh < 0 || hh % a.sH || hh / a.sH >= a.HO) continue;
                ho_off = (hh / a.sH) * a.WO;
              }
              for (int kw = kw_lo; kw <= kw_hi; kw++) {
                int px;
                if (unit) {
                  px = base0 - kh * a.WO - kw;
                } else {
                  const int ww = w + a.padW - kw * a.dW;
                  if (ww < 0 || ww % a.sW || ww / a.sW >= a.WO) continue;
                  px = il * HWo + ho_off + ww / a.sW - q_lo;
                }
                if (whole || ((unsigned)px < (unsigned)span)) {
                  const float* src = cb + (kh * a.kW + kw) * a.colL + px;
#pragma unroll
                  for (int c = 0; c < 8; c++)
                    if (c < nc) sum[c] = __fadd_rn(sum[c], src[c * cstride]);
                }
              }
            }
#pragma unroll
            for (int c = 0; c < 8; c++) {
              if (c < nc) {
                float* ga = img_p + (c0 + c) * a.ipg * HWi + o;
                *ga = __fadd_rn(*ga, sum[c]);
              }
            }
          }
        }
        bar_sync_named(1, 512);                                 // col_s may be overwritten by the next tile; gin_s complete
        w_p2 += pclk() - tp2;
      }
      const long long tfl = pclk();
      // flush the finished images (coalesced), leaving zeros behind for the next group; the first barrier of the next
      // tile orders these accesses before its accumulation phase
      const int nflush = pre ? 0 : cpc_here * nout;
      for (int i = tid; i < nflush; i += 512) {
        const int cl = i / nout, o = i - cl * nout;
        const int il = o / HWi, rem = o - il * HWi;
        float* ga = img_p + cl * a.ipg * HWi + o;
        const float x = *ga;
        *ga = 0.f;
        a.gin[((n0 + il) * a.C + ci0 + cl) * (int64_t)HWi + rem] = x;
      }
      w_fl += pclk() - tfl;
    }
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 256) { a.dbg[5] = pclk() - tstart; a.dbg[6] = w_df; a.dbg[7] = w_p1; a.dbg[8] = w_p2; a.dbg[9] = w_fl; }
  }

  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<1>(tmem_base, 512);
}

typedef CUresult (*PFN_TmapEncodeTiled3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// grad_input = col2im(W^T * grad_output); *done == false -> the shape does not fit this path (caller falls back)
int conv2d_dgrad_col2im_tc_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* grad_output,
                               const float* kernel, float* grad_input, bool only_if_fast, bool* done) {
  *done = false;
  if (!gemm_f32_tc_available()) return AM_OK;
  const int KK = (int)(d.kH * d.kW);
  if (d.Cout > 64 || KK > 64 || KK < 1) return AM_OK;
  if (d.H * d.W >= (1 << 20) || Ho * Wo >= (1 << 20) || d.C >= (1 << 20)) return AM_OK;
  const int HWi = (int)(d.H * d.W), HWo = (int)(Ho * Wo);
  DgradTcArgs a{};
  a.gout = grad_output; a.gin = grad_input; a.N = d.N;
  a.C = (int)d.C; a.H = (int)d.H; a.W = (int)d.W; a.CO = (int)d.Cout; a.HO = (int)Ho; a.WO = (int)Wo;
  a.kH = (int)d.kH; a.kW = (int)d.kW; a.padH = (int)d.padH; a.padW = (int)d.padW; a.sH = (int)d.strideH; a.sW = (int)d.strideW;
  a.dH = (int)d.dilH; a.dW = (int)d.dilW;
  a.Kpad = d.Cout <= 32 ? 32 : 64;
  a.cpc = 128 / KK;
  if (a.cpc > a.C) a.cpc = a.C;
  a.nchunks = (a.C + a.cpc - 1) / a.cpc;
  a.ipg = HWo <= 128 ? 128 / HWo : 1;
  if (a.ipg > 16) a.ipg = 16;
  a.tpg = (a.ipg * HWo + 127) / 128;
  a.ngroups = ceil_div(d.N, a.ipg);
  const bool unit = d.strideH == 1 && d.strideW == 1 && d.dilH == 1 && d.dilW == 1;
  a.WP = a.WO + a.kW - 1;
  a.padded = (a.tpg == 1 && unit && a.kW <= 8 && (a.kW - 1) + a.ipg * a.HO * a.WP <= kPadColL) ? 1 : 0;
  a.colL = a.padded ? kPadColL : 128;
  // measured (profiles/r01_bringup.md): the zero-padded single-tile form beats the direct SIMT kernel once the GEMM has
  // some width (C*kH*kW >= 128 rows); the multi-tile form (large images) does not yet
  if (only_if_fast && !(a.padded && (int64_t)a.C * KK >= 128)) return AM_OK;
  const int sms = sm_count();
  if (a.nchunks > sms) return AM_OK;
  const size_t plane = (size_t)(a.Kpad / 32) * 16384;
  const size_t smem = 1024 + 2 * plane + (size_t)a.cpc * KK * a.colL * 4 + ((size_t)a.cpc * a.ipg * HWi + 4) * 4 + 16 + 8 * 10 + 64;
  if (smem > 227 * 1024) return AM_OK;

  // W^T chunks, pre-split into tf32 hi / lo planes
  void* ws = nullptr;
  const size_t wplane = (size_t)a.nchunks * 128 * a.Kpad * sizeof(float);
  int rc = workspace(kWsConvW, 2 * wplane + 256, &ws);
  if (rc) return rc;
  float* whi = (float*)ws; float* wlo = (float*)((char*)ws + wplane);
  dgrad_tc_pack_weights_kernel<<<(unsigned)ceil_div((int64_t)a.nchunks * 128 * a.Kpad, 256), 256, 0, st>>>(
      kernel, whi, wlo, a.C, a.CO, KK, a.cpc, a.nchunks, a.Kpad);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());

  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  AM_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) return AM_OK;
  auto encode = (PFN_TmapEncodeTiled3)fn;
  CUtensorMap tms[2];
  float* planes[2] = {whi, wlo};
  for (int i = 0; i < 2; i++) {
    cuuint64_t gdim[2] = {(cuuint64_t)a.Kpad, (cuuint64_t)a.nchunks * 128};
    cuuint64_t gstride[1] = {(cuuint64_t)a.Kpad * sizeof(float)};
    cuuint32_t box[2] = {32u, 128u};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = encode(&tms[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, planes[i], gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("conv dgrad tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return AM_ERR_CUDA; }
  }
  int dbg_env;
  dbg_env = tuning(kTuneConvTcDebug) & 1;
  a.dbg = nullptr;
  if (dbg_env) {
    void* base = nullptr;
    if ((rc = workspace(kWsMisc, 64 * sizeof(int) + 1024, &base))) return rc;
    a.dbg = (long long*)((char*)base + 512);
    AM_CUDA_TRY(cudaMemsetAsync(a.dbg, 0, 16 * sizeof(long long), st));
  }
  int grid = (sms / a.nchunks) * a.nchunks;
  if ((int64_t)grid > a.ngroups * a.nchunks) grid = (int)(a.ngroups * a.nchunks);
  AM_CUDA_TRY(cudaFuncSetAttribute(conv_dgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv_dgrad_tc_kernel<<<grid, 768, smem, st>>>(tms[0], tms[1], a);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  if (a.dbg) {
    long long h[16];
    AM_CUDA_TRY(cudaMemcpyAsync(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    AM_CUDA_TRY(cudaStreamSynchronize(st));
    fprintf(stderr, "[dgrad_tc dbg] grid=%d chunks=%d cpc=%d ipg=%d tpg=%d groups/cta~%lld | mma: total %lld wait_d_empty %lld wait_a_full %lld | operand: total %lld wait_a_empty %lld | col2im: total %lld wait_d_full %lld phase1 %lld phase2 %lld flush %lld\n",
            grid, a.nchunks, a.cpc, a.ipg, a.tpg, (long long)(a.ngroups * a.nchunks / grid), h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9]);
  }
  *done = true;
  return AM_OK;
}


// =====================================================================================================================
// Weight gradient on the tensor cores: the forward implicit GEMM with the roles of "pixels" and "k" exchanged,
//
//   dW[R, co] = sum over images n and output pixels q of  col[n, R, q] * gout[n, co, q],   R = (ci, kh, kw)  (M, TMEM lanes)
//
// (conv.nim:140  gW += gout[i] * col^T, summed over the batch).  Row R = Kc is a row of ones, so column... row Kc of the
// result is grad_bias (nnp_convolution.nim:91-94).  A CTA owns one chunk of 128 rows R and a slice of the batch; per
// stage of 32 pixels the gather groups expand the staged input channels into the A operand IN TENSOR MEMORY
// (lane = row R, column = pixel), the loader warps split the matching gout block [co x 32 pixels] into tf32 hi / lo and
// write it K-major / 128B-swizzled into shared memory (B operand), and one thread issues 12 tcgen05.mma.  Chains of
// `flush_st` stages go to one of two TMEM accumulators and are drained into fp32 registers with round-to-nearest adds
// (the tensor core truncates when it accumulates).  Partial sums per batch slice go to part[slice][co][Nv]; the
// existing fixed-order wgrad_reduce_kernel adds the slices (deterministic).
struct WgradTcArgs {
  const float* x;       // [N][C][H][W]
  const float* gout;    // [N][CO][HO][WO]
  float* part;          // [nslices][CO][Nv]
  int64_t N;
  int C, H, W, CO, HO, WO, kH, kW, padH, padW, sH, sW, dH, dW;
  int Kc, Nv;           // C*kH*kW, Kc + 1 (the ones row)
  int NP;               // Cout padded to 16 / 32 / 64: UMMA N
  int nchunks, nslices;
  int spi;              // stages of 32 pixels per image
  int flush_st;         // stages per accumulation chain
  int SB;               // depth of the gout ring
  int groups;           // gather groups (of 128 threads), <= 3
  uint32_t raw_bytes;   // bytes of one staged-input buffer
  int nraw_log2;        // log2 of the number of staged-input buffers (2, 4 or 8)
  int checked;          // the convolution has padding: bounds test per gathered element
  int vec;              // gout rows can be read with 128-bit loads
  long long* dbg;       // optional per-role cycle counters of CTA 0 (AM_CONVTC_DEBUG=1)
  int skip;             // experiments (knob convtc_debug >> 1): bit 0 = gather warps skip their loads / TMEM stores, bit 1 = loaders skip split / stores
  int b_tma;            // grad_output stages arrive by TMA from pre-split hi / lo planes (no loader warps)
};

constexpr int kWgAStages = 6;

int make_tmap_f32_nd_sw128(CUtensorMap* tm, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box);     // gemm_f32_tc.cu

// grad_output -> tf32 hi / lo planes, once per call (storing them from the data-gradient kernel, which splits every element
// for its own A operand, was measured: 100 extra stores per thread and tile on its operand warps cost 150 us, this pass 25): the 4 chunk CTAs of a batch slice then fetch their stages by TMA instead
// of each splitting the same block with its loader warps (which were the slowest role of the kernel)
__global__ void __launch_bounds__(256) wgrad_split_gout_kernel(const float4* __restrict__ g, float4* __restrict__ hi,
                                                              float4* __restrict__ lo, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(g + i);
    const float x[4] = {v.x, v.y, v.z, v.w};
    float h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      uint32_t hb, lb;
      ptx::split_tf32(x[e], hb, lb);
      h[e] = __uint_as_float(hb); l[e] = __uint_as_float(lb);
    }
    hi[i] = make_float4(h[0], h[1], h[2], h[3]);
    lo[i] = make_float4(l[0], l[1], l[2], l[3]);
  }
}
#define WG_TWAIT(counter, bar, ph) do { const long long t0_ = pclk(); ptx::mbar_wait(bar, ph); counter += pclk() - t0_; } while (0)


__global__ void __launch_bounds__(768, 1) conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmGhi,
                                                               const __grid_constant__ CUtensorMap tmGlo, const WgradTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)a.NP * 128u;                     // one plane of a gout stage: NP rows x 32 pixels
  const uint32_t stage_bytes = 2u * b_bytes;
  const int SB = a.SB;
  const uint32_t raw_base = smem_base + (uint32_t)SB * stage_bytes;
  // staged-input ring: an image is only `spi` stages (~0.5 us of MMAs) of work but its bulk copy takes ~1.5 us to land: with
  // two buffers (one image of prefetch) the whole kernel ran at the latency of one bulk copy per image — 170 us with the
  // gather and loader work switched off (profiles/r02_conv_tc_issue_analysis.md).  Up to 8 buffers = 7 images ahead.
  const int NRAW = 1 << a.nraw_log2;
  const uint32_t bar_base = raw_base + (uint32_t)NRAW * a.raw_bytes;
  auto full_a = [&](int s) { return bar_base + 8u * s; };               // 6
  auto empty_a = [&](int s) { return bar_base + 8u * (6 + s); };        // 6
  auto full_b = [&](int s) { return bar_base + 8u * (12 + s); };        // 8
  auto empty_b = [&](int s) { return bar_base + 8u * (20 + s); };       // 8
  auto tfull_bar = [&](int b) { return bar_base + 8u * (28 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (30 + b); };
  auto raw_full = [&](int b) { return bar_base + 8u * (40 + b); };       // 8
  auto raw_empty = [&](int b) { return bar_base + 8u * (48 + b); };      // 8
  const uint32_t tmem_slot = bar_base + 8u * 36;
  const uint32_t ptab = bar_base + 8u * 56;                               // int2 per output pixel: {byte offset of (h0, w0), h0 | w0 << 16}

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = ptx::lane_id();
  const int G = a.groups;
  const int KK = a.kH * a.kW, HWi = a.H * a.W, HWo = a.HO * a.WO;
  const int chunk = (int)blockIdx.x % a.nchunks, slice = (int)blockIdx.x / a.nchunks;
  for (int q = threadIdx.x; q < a.spi * 32; q += blockDim.x) {            // pixel walk of an image, shared by all rows
    const int ho = q / a.WO, wo = q - ho * a.WO;
    const int h0 = ho * a.sH - a.padH, w0 = wo * a.sW - a.padW;
    const int e0 = (q < HWo) ? (h0 * a.W + w0) * 4 : 0, e1 = (q < HWo) ? ((h0 & 0xffff) | (w0 << 16)) : 0x7fff7fff;
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(ptab + 8u * q), "r"(e0), "r"(e1) : "memory");
  }
  const int R0 = chunk * 128;
  const int64_t nimg = (a.N > slice) ? (a.N - slice + a.nslices - 1) / a.nslices : 0;    // images slice, slice + nslices, ...
  const int64_t T = nimg * a.spi;                                         // stages this CTA walks through
  // input channels this chunk's rows touch
  const int Rl = (R0 < a.Kc) ? R0 : a.Kc - 1, Rh = (R0 + 127 < a.Kc) ? R0 + 127 : a.Kc - 1;
  const int ci_first = Rl / KK, nci = Rh / KK - ci_first + 1;

  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < kWgAStages; s++) { ptx::mbar_init(full_a(s), 128); ptx::mbar_init(empty_a(s), 1); }
    for (int s = 0; s < SB; s++) { ptx::mbar_init(full_b(s), a.b_tma ? 1 : 128); ptx::mbar_init(empty_b(s), 1); }
    for (int b = 0; b < 2; b++) { ptx::mbar_init(tfull_bar(b), 1); ptx::mbar_init(tempty_bar(b), 4); }
    for (int b = 0; b < NRAW; b++) { ptx::mbar_init(raw_full(b), 1); ptx::mbar_init(raw_empty(b), 128u * (uint32_t)G); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<1>(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tmem_a0 = tmem_base + 128u;                              // 6 operand stages of 64 columns after the accumulators

  // register budget (24 warps at 80): control 56, gather / loader warps 72, accumulate warps 136
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ===================== staged-input producer: the chunk's channels of one image per bulk copy =====================
      if (ptx::elect_one()) {
        const uint32_t bytes = (uint32_t)(nci * HWi) * 4u;
        for (int64_t i = 0; i < nimg; i++) {
          const int b = (int)(i & (NRAW - 1));
          ptx::mbar_wait(raw_empty(b), (uint32_t)((i >> a.nraw_log2) & 1) ^ 1u);
          const int64_t n = slice + i * a.nslices;
          ptx::mbar_arrive_expect_tx(raw_full(b), bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(raw_base + (uint32_t)b * a.raw_bytes), "l"(a.x + (n * a.C + ci_first) * (int64_t)HWi), "r"(bytes), "r"(raw_full(b)) : "memory");
        }
      }
    } else if (warp == 1) {
      // ===================== UMMA issuer =====================
      if (ptx::elect_one()) {
        const uint64_t dhi = ptx::umma_desc_hi(1024, 128);
        const uint32_t idesc = ptx::umma_idesc_tf32(128, (uint32_t)a.NP);
        uint32_t chain = 0;
        long long w_te = 0, w_fa = 0, w_fb = 0;
        const long long tstart = pclk();
        // ring positions, wait parities and the position inside the accumulation chain advance by compare-and-wrap (the
        // 64-bit `it % flush`, `it % SB`, `it / SB` of the first version were ~150 dependent instructions per stage in
        // the single issuing thread)
        int in_chain = 0, sa = 0, sbi = 0; uint32_t pa = 0u, pb = 0u;
        uint32_t a_hi0 = tmem_a0;
        const uint64_t b_first = ptx::umma_desc(dhi, smem_base), b_step = (uint64_t)(stage_bytes >> 4), b_lo_off = (uint64_t)(b_bytes >> 4);
        uint64_t b0 = b_first;
        for (int64_t it = 0; it < T; it++) {
          const int buf = chain & 1;
          if (in_chain == 0) {
            WG_TWAIT(w_te, tempty_bar(buf), ((chain >> 1) & 1u) ^ 1u);
            ptx::tc_fence_after();
          }
          WG_TWAIT(w_fa, full_a(sa), pa);
          WG_TWAIT(w_fb, full_b(sbi), pb);
          ptx::tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)buf * 64u;
          const uint32_t a_lo0 = a_hi0 + 32u;
#pragma unroll
          for (int k8 = 0; k8 < 4; k8++) {
            const uint64_t b_hi = b0 + 2u * k8, b_lo = b0 + b_lo_off + 2u * k8;   // descriptor address field: 16-byte units
            ptx::umma_tf32_ts(d, a_lo0 + 8u * k8, b_hi, idesc, (in_chain | k8) ? 1u : 0u);
            ptx::umma_tf32_ts(d, a_hi0 + 8u * k8, b_lo, idesc, 1u);
            ptx::umma_tf32_ts(d, a_hi0 + 8u * k8, b_hi, idesc, 1u);
          }
          ptx::umma_commit<1>(empty_a(sa));
          ptx::umma_commit<1>(empty_b(sbi));
          if (in_chain == a.flush_st - 1 || it == T - 1) { ptx::umma_commit<1>(tfull_bar(buf)); chain++; in_chain = 0; }
          else in_chain++;
          a_hi0 += 64u;
          if (++sa == kWgAStages) { sa = 0; pa ^= 1u; a_hi0 = tmem_a0; }
          b0 += b_step;
          if (++sbi == SB) { sbi = 0; pb ^= 1u; b0 = b_first; }
        }
        if (a.dbg && blockIdx.x == 0) { a.dbg[0] = pclk() - tstart; a.dbg[1] = w_te; a.dbg[2] = w_fa; a.dbg[3] = w_fb; }
      }
    } else if (warp == 2 && a.b_tma) {
      // ===================== grad_output producer (TMA): [NP rows x 32 pixels] boxes of the pre-split hi / lo planes,
      // 128-byte swizzled by the TMA unit; the four warps that used to load and split them are a fourth gather group
      if (ptx::elect_one()) {
        ptx::prefetch_tensormap(&tmGhi); ptx::prefetch_tensormap(&tmGlo);
        int sbi = 0; uint32_t pb = 1u;
        int s_run = 0; int64_t n_run = slice;
        for (int64_t it = 0; it < T; it++) {
          ptx::mbar_wait(empty_b(sbi), pb);
          const uint32_t sb = smem_base + (uint32_t)sbi * stage_bytes;
          ptx::mbar_arrive_expect_tx(full_b(sbi), stage_bytes);
          ptx::tma_load_3d(sb, &tmGhi, full_b(sbi), s_run * 32, 0, (int)n_run);
          ptx::tma_load_3d(sb + b_bytes, &tmGlo, full_b(sbi), s_run * 32, 0, (int)n_run);
          if (++s_run == a.spi) { s_run = 0; n_run += a.nslices; }
          if (++sbi == SB) { sbi = 0; pb ^= 1u; }
        }
      }
    }
  } else if (warp < 4 + 4 * G) {
    // ===================== gather groups: A[R, pixel] -> tensor memory =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int g = (warp - 4) >> 2;
    const int r = (int)(threadIdx.x - 128) & 127;
    const int R = R0 + r;
    const uint32_t t_lane = tmem_a0 + ((uint32_t)((warp & 3) * 32) << 16);
    int kind = 0, base = 0, khd = 0, kwd = 0;            // 0: zero row, 1: tap row, 2: ones row (grad_bias)
    if (R < a.Kc) {
      const int ci = R / KK, tap = R - ci * KK, kh = tap / a.kW, kw = tap - kh * a.kW;
      kind = 1; khd = kh * a.dH; kwd = kw * a.dW;
      base = (ci - ci_first) * HWi + khd * a.W + kwd;
    } else if (R == a.Kc) {
      kind = 2;
    }
    long long w_rf = 0, w_ea = 0;
    const long long tstart = pclk();
    // stage counters advance by compare-and-wrap (no 64-bit `it % G`, `it % stages`, `it / stages` per stage and thread)
    int turn = 0, sa_run = 0; uint32_t pa_run = 1u;              // whose stage this is | ring position | wait parity of the ring lap
    for (int64_t i = 0; i < nimg; i++) {
      const int b = (int)(i & (NRAW - 1));
      WG_TWAIT(w_rf, raw_full(b), (uint32_t)((i >> a.nraw_log2) & 1));      // every group waits for every image (keeps the phases of raw_empty in step)
      const uint32_t xb = raw_base + (uint32_t)b * a.raw_bytes + (uint32_t)(base * 4);
      for (int s = 0; s < a.spi; s++) {
        const int sa = sa_run; const uint32_t pa = pa_run;
        const bool mine = turn == g;
        if (++turn == G) turn = 0;
        if (++sa_run == kWgAStages) { sa_run = 0; pa_run ^= 1u; }
        if (!mine) continue;
        WG_TWAIT(w_ea, empty_a(sa), pa);
        ptx::tc_fence_after();
        const uint32_t ta = t_lane + 64u * (uint32_t)sa;
        const int q0 = s * 32;
#pragma unroll
        for (int half = 0; half < 2; half++) {
          if (a.skip & 1) break;
          float v[16];
          if (kind == 1) {
#pragma unroll
            for (int j2 = 0; j2 < 8; j2++) {                   // two pixel-table entries per 128-bit broadcast load
              int e0x, e0y, e1x, e1y;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(e0x), "=r"(e0y), "=r"(e1x), "=r"(e1y)
                           : "r"(ptab + 8u * (uint32_t)(q0 + half * 16 + 2 * j2)));
              float x0 = 0.f, x1 = 0.f;
              if (a.checked) {
                // the table holds h0 / w0 as 16-bit fields (0x7fff for pixels past the image: never in range)
                const int ha = (int)(short)(e0y & 0xffff) + khd, wa = (e0y >> 16) + kwd;
                const int hb = (int)(short)(e1y & 0xffff) + khd, wb = (e1y >> 16) + kwd;
                if ((unsigned)ha < (unsigned)a.H && (unsigned)wa < (unsigned)a.W) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x0) : "r"(xb + (uint32_t)e0x));
                if ((unsigned)hb < (unsigned)a.H && (unsigned)wb < (unsigned)a.W) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x1) : "r"(xb + (uint32_t)e1x));
              } else {
                if (q0 + half * 16 + 2 * j2 < HWo) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x0) : "r"(xb + (uint32_t)e0x));
                if (q0 + half * 16 + 2 * j2 + 1 < HWo) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x1) : "r"(xb + (uint32_t)e1x));
              }
              v[2 * j2] = x0; v[2 * j2 + 1] = x1;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = (kind == 2 && q0 + half * 16 + j < HWo) ? 1.f : 0.f;
          }
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; j++) {
            ptx::split_tf32(v[j], hi[j], lo[j]);
          }
          ptx::tmem_st_32x16(ta + 16u * half, hi);
          ptx::tmem_st_32x16(ta + 32u + 16u * half, lo);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_a(sa)) : "memory");
      }
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(raw_empty(b)) : "memory");
    }
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 128) { a.dbg[4] = pclk() - tstart; a.dbg[5] = w_rf; a.dbg[6] = w_ea; }
  } else if (warp < 16 || (a.b_tma && warp < 20)) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");       // unused gather slots (and the loader slot when TMA feeds B)
  } else if (warp < 20) {
    // ===================== loader warps: gout block [co x 32 pixels] -> tf32 hi/lo, K-major swizzled smem =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (!a.b_tma) {
    const int t = (int)threadIdx.x - 512;                       // 0..127
    const int tpr = 128 / a.NP;                                 // threads per row: 2, 4 or 8
    const int co = t / tpr, part = t - co * tpr;
    const int ppt = 32 / tpr;                                   // pixels per thread: 16, 8 or 4
    const uint32_t row_off = (uint32_t)co * 128u, sw = (uint32_t)(co & 7);
    const int nv = ppt / 4;                                     // 128-bit pieces per thread and stage: 1, 2 or 4
    // global loads of stage it + 1 are in flight while stage it is split and stored (the loader is otherwise
    // latency-bound: one dependent L2 round trip per stage)
    auto fetch = [&](int64_t it, float4 (&dst)[4]) {
      const int64_t i = (int64_t)((uint32_t)it / (uint32_t)a.spi);      // T < 2^31 stages: 32-bit division
      const int s = (int)(it - i * a.spi);
      const int64_t n = slice + i * a.nslices;
      const int q0 = s * 32 + part * ppt;
      const float* src = a.gout + (n * a.CO + co) * (int64_t)HWo + q0;
#pragma unroll
      for (int v4 = 0; v4 < 4; v4++) {
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        const int q = q0 + 4 * v4;
        if (v4 < nv && co < a.CO) {
          if (a.vec) {
            if (q < HWo) f = __ldg(reinterpret_cast<const float4*>(src + 4 * v4));
          } else {
            if (q < HWo) f.x = __ldg(src + 4 * v4);
            if (q + 1 < HWo) f.y = __ldg(src + 4 * v4 + 1);
            if (q + 2 < HWo) f.z = __ldg(src + 4 * v4 + 2);
            if (q + 3 < HWo) f.w = __ldg(src + 4 * v4 + 3);
          }
        }
        dst[v4] = f;
      }
    };
    // a stage is consumed in ~500 cycles (12 MMAs) but one L2 round trip takes longer: the loads of stage it + 2 are issued
    // before stage it is split and stored (three register buffers, rotated by a 3x unrolled loop — no register copies).
    // With a distance of one stage the MMA thread spent 37 % of its time waiting for grad_output (r02_convtc_waits.txt).
    // Fast path (128-bit aligned rows, every stage full: HWo % 32 == 0): straight-line loads through a running pointer.  The
    // generic fetch above costs ~250 instructions of branches and a division per stage and made the 4 loader warps — which
    // run every stage one after the other — the slowest role of the kernel (the MMA thread waited 60 % of its time for
    // grad_output stages even with the split and the stores switched off, profiles/r02_conv_tc_issue_analysis.md).
    const bool fast = a.vec && (HWo % 32 == 0);
    const float* run_src = a.gout + ((int64_t)slice * a.CO + co) * (int64_t)HWo + part * ppt;
    const int64_t img_step = (int64_t)a.nslices * a.CO * HWo - (int64_t)(a.spi - 1) * 32;
    int run_s = 0;
    const bool row_ok = co < a.CO;
    auto fetch_seq = [&](int64_t it, float4 (&dst)[4]) {            // called with it = 0, 1, 2, ... in order
      if (!fast) { fetch(it, dst); return; }
#pragma unroll
      for (int v4 = 0; v4 < 4; v4++)
        dst[v4] = (row_ok && v4 < nv) ? __ldg(reinterpret_cast<const float4*>(run_src) + v4) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (++run_s == a.spi) { run_s = 0; run_src += img_step; } else run_src += 32;
    };
    float4 bufA[4], bufB[4], bufC[4];
    if (T > 0) fetch_seq(0, bufA);
    if (T > 1) fetch_seq(1, bufB);
    long long w_eb = 0;
    const long long tstart = pclk();
    int sbi = 0; uint32_t pb = 1u;
    auto consume = [&](const float4 (&cur)[4]) {
      WG_TWAIT(w_eb, empty_b(sbi), pb);
      const uint32_t sb = smem_base + (uint32_t)sbi * stage_bytes;
#pragma unroll
      for (int v4 = 0; v4 < 4; v4++) {
        if (v4 < nv && !(a.skip & 2)) {
          const float v[4] = {cur[v4].x, cur[v4].y, cur[v4].z, cur[v4].w};
          float h4[4], l4[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            uint32_t hb, lb;
            ptx::split_tf32(v[e], hb, lb);
            h4[e] = __uint_as_float(hb); l4[e] = __uint_as_float(lb);
          }
          const uint32_t c16 = (uint32_t)(part * nv + v4);                  // 16-byte chunk of the 128-byte row
          const uint32_t off = row_off + ((c16 ^ sw) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb + off), "f"(h4[0]), "f"(h4[1]), "f"(h4[2]), "f"(h4[3]) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb + b_bytes + off), "f"(l4[0]), "f"(l4[1]), "f"(l4[2]), "f"(l4[3]) : "memory");
        }
      }
      if (!(a.skip & 4)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_b(sbi)) : "memory");
      if (++sbi == SB) { sbi = 0; pb ^= 1u; }
    };
    for (int64_t it = 0; it < T; it += 3) {
      if (it + 2 < T) fetch_seq(it + 2, bufC);
      consume(bufA);
      if (it + 1 < T) {
        if (it + 3 < T) fetch_seq(it + 3, bufA);
        consume(bufB);
      }
      if (it + 2 < T) {
        if (it + 4 < T) fetch_seq(it + 4, bufB);
        consume(bufC);
      }
    }
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 512) { a.dbg[7] = pclk() - tstart; a.dbg[8] = w_eb; }
    }
  } else {
    // ===================== accumulate warps =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
    const int qd = warp & 3;
    const int r = qd * 32 + (int)lane;
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; i++) acc[i] = 0.f;
    const int64_t nchains = (T + a.flush_st - 1) / a.flush_st;
    long long w_tf = 0;
    const long long tstart = pclk();
    for (int64_t c = 0; c < nchains; c++) {
      const int buf = (int)(c & 1);
      WG_TWAIT(w_tf, tfull_bar(buf), (uint32_t)((c >> 1) & 1));
      ptx::tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)buf * 64u;
      if (a.NP > 32) {                                       // both halves in flight before the one wait
        uint32_t r0[32], r1[32];
        ptx::tmem_ld_32x32(t0, r0);
        ptx::tmem_ld_32x32(t0 + 32, r1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) acc[i] = __fadd_rn(acc[i], __uint_as_float(r0[i]));
#pragma unroll
        for (int i = 0; i < 32; i++) acc[32 + i] = __fadd_rn(acc[32 + i], __uint_as_float(r1[i]));
      } else {
        uint32_t r0[32];
        ptx::tmem_ld_32x32(t0, r0);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) acc[i] = __fadd_rn(acc[i], __uint_as_float(r0[i]));
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(buf)) : "memory");
    }
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 640) { a.dbg[9] = pclk() - tstart; a.dbg[10] = w_tf; }
    const int R = R0 + r;
    if (R < a.Nv) {
      float* dst = a.part + (int64_t)slice * a.CO * a.Nv + R;
#pragma unroll
      for (int co = 0; co < 64; co++)
        if (co < a.CO) dst[(int64_t)co * a.Nv] = acc[co];
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<1>(tmem_base, 512);
}

// partial weight / bias gradients per batch slice on the tensor cores; *done == false -> shape does not fit
int conv2d_wgrad_tc_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* input,
                        const float* grad_output, float** part_out, int* groups_out, bool* done) {
  *done = false;
  if (!gemm_f32_tc_available()) return AM_OK;
  const int KK = (int)(d.kH * d.kW);
  if (d.Cout > 64 || d.Cout < 1 || KK < 1) return AM_OK;
  if (d.H * d.W >= (1 << 20) || Ho * Wo >= (1 << 20) || d.C * (int64_t)KK >= (1 << 24)) return AM_OK;
  if (d.H >= 32000 || d.W >= 32000) return AM_OK;                          // 16-bit fields of the pixel table
  const int HWi = (int)(d.H * d.W), HWo = (int)(Ho * Wo);
  if (HWi % 4 != 0 || (reinterpret_cast<uintptr_t>(input) & 15) != 0) return AM_OK;     // 16-byte bulk copies of whole channels
  WgradTcArgs a{};
  a.x = input; a.gout = grad_output; a.N = d.N;
  a.C = (int)d.C; a.H = (int)d.H; a.W = (int)d.W; a.CO = (int)d.Cout; a.HO = (int)Ho; a.WO = (int)Wo;
  a.kH = (int)d.kH; a.kW = (int)d.kW; a.padH = (int)d.padH; a.padW = (int)d.padW; a.sH = (int)d.strideH; a.sW = (int)d.strideW;
  a.dH = (int)d.dilH; a.dW = (int)d.dilW;
  a.Kc = a.C * KK; a.Nv = a.Kc + 1;
  a.NP = a.CO <= 16 ? 16 : (a.CO <= 32 ? 32 : 64);
  a.nchunks = (a.Nv + 127) / 128;
  const int sms = sm_count();
  if (a.nchunks > sms) return AM_OK;
  a.nslices = sms / a.nchunks;
  if ((int64_t)a.nslices > d.N) a.nslices = (int)d.N;
  if (a.nslices < 1) a.nslices = 1;
  a.spi = (HWo + 31) / 32;
  int flush_env, groups_env;
  flush_env = tuning(kTuneTcFlushKb) > 0 ? tuning(kTuneTcFlushKb) : 2;
  groups_env = (tuning(kTuneConvTcGroups) >= 1 && tuning(kTuneConvTcGroups) <= 3) ? tuning(kTuneConvTcGroups) : 3;
  a.flush_st = flush_env;
  a.groups = groups_env > 3 ? 3 : groups_env;              // a fourth group takes the loader warps when TMA feeds grad_output (below)
  a.checked = (a.padH != 0 || a.padW != 0) ? 1 : 0;
  a.vec = (HWo % 4 == 0 && (reinterpret_cast<uintptr_t>(grad_output) & 15) == 0) ? 1 : 0;
  const int nci_max = 127 / KK + 2 < a.C ? 127 / KK + 2 : a.C;
  a.raw_bytes = (uint32_t)round_up((int64_t)nci_max * HWi * 4, 128);
  const size_t stage_bytes = 2 * (size_t)a.NP * 128;
  a.nraw_log2 = 1;
  for (int lg = 3; lg >= 2; lg--)       // as many staged-input buffers as leave room for a grad_output ring of 6 stages
    if (1024 + ((size_t)a.raw_bytes << lg) + 8 * 56 + (size_t)a.spi * 32 * 8 + 64 + 6 * stage_bytes <= 227 * 1024) { a.nraw_log2 = lg; break; }
  const size_t fixed = 1024 + ((size_t)a.raw_bytes << a.nraw_log2) + 8 * 56 + (size_t)a.spi * 32 * 8 + 64;
  if (fixed + 2 * stage_bytes > 227 * 1024) return AM_OK;
  int SB = (int)((227 * 1024 - fixed) / stage_bytes);
  if (SB > 8) SB = 8;
  a.SB = SB;
  void* part = nullptr;
  int rc = workspace(kWsConv, (size_t)a.nslices * a.CO * a.Nv * sizeof(float), &part);
  if (rc) return rc;
  a.part = (float*)part;
  const size_t smem = fixed + (size_t)SB * stage_bytes;
  AM_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  a.skip = tuning(kTuneConvTcDebug) >> 1;
  int dbg_env;
  dbg_env = tuning(kTuneConvTcDebug) & 1;
  a.dbg = nullptr;
  if (dbg_env) {
    void* base = nullptr;
    if ((rc = workspace(kWsMisc, 64 * sizeof(int) + 1024, &base))) return rc;
    a.dbg = (long long*)((char*)base + 512);
    AM_CUDA_TRY(cudaMemsetAsync(a.dbg, 0, 16 * sizeof(long long), st));
  }
  // pre-split grad_output planes + tensor maps (rows of 16-byte multiples, aligned base); otherwise the loader warps do it
  CUtensorMap tmG[2];
  memset(tmG, 0, sizeof(tmG));
  a.b_tma = 0;
  int64_t gtotal = 0;
  if (gout_planes_ok(d, HWo, grad_output, &gtotal)) {
    void* planes = nullptr;
    if ((rc = workspace(kWsGoutSplit, (size_t)gtotal * 2 * sizeof(float) + 256, &planes))) return rc;
    float* ghi = (float*)planes; float* glo = ghi + gtotal;
    int64_t blocks = ceil_div(gtotal / 4, 256);
    if (blocks > 16 * (int64_t)sm_count()) blocks = 16 * (int64_t)sm_count();
    wgrad_split_gout_kernel<<<(unsigned)blocks, 256, 0, st>>>((const float4*)grad_output, (float4*)ghi, (float4*)glo, gtotal / 4);
    g_launch_count++;
    AM_CUDA_TRY(cudaGetLastError());
    const uint64_t dims[3] = {(uint64_t)HWo, (uint64_t)a.CO, (uint64_t)a.N};
    const uint64_t strides[2] = {(uint64_t)HWo * 4, (uint64_t)HWo * a.CO * 4};
    const uint32_t box[3] = {32u, (uint32_t)a.NP, 1u};
    if ((rc = make_tmap_f32_nd_sw128(&tmG[0], ghi, 3, dims, strides, box))) return rc;
    if ((rc = make_tmap_f32_nd_sw128(&tmG[1], glo, 3, dims, strides, box))) return rc;
    a.b_tma = 1;
    if (tuning(kTuneConvTcGroups) == 0 || tuning(kTuneConvTcGroups) == 4) a.groups = 4;
  }
  conv_wgrad_tc_kernel<<<a.nchunks * a.nslices, 768, smem, st>>>(tmG[0], tmG[1], a);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  if (a.dbg) {
    long long h[16];
    AM_CUDA_TRY(cudaMemcpyAsync(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    AM_CUDA_TRY(cudaStreamSynchronize(st));
    fprintf(stderr, "[wgrad_tc dbg] chunks=%d slices=%d spi=%d SB=%d NP=%d | mma: total %lld wait_tempty %lld wait_full_a %lld wait_full_b %lld | gather0: total %lld wait_raw_full %lld wait_empty_a %lld | loader: total %lld wait_empty_b %lld | acc: total %lld wait_tfull %lld\n",
            a.nchunks, a.nslices, a.spi, a.SB, a.NP, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10]);
  }
  *part_out = a.part;
  *groups_out = a.nslices;
  *done = true;
  return AM_OK;
}

}  // namespace am
