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
      const long long tstart = clock64();
      for (int64_t g = g0; g < a.ngroups; g += gstep) {
        for (int t = 0; t < a.tpg; t++, it++) {
          const int s = it & 1;
          const uint32_t ph = (it >> 1) & 1u;
          { const long long t0_ = clock64(); ptx::mbar_wait(d_empty(s), ph ^ 1u); w_de += clock64() - t0_; }
          { const long long t0_ = clock64(); ptx::mbar_wait(a_full(s), ph); w_af += clock64() - t0_; }
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
      if (a.dbg && blockIdx.x == 0) { a.dbg[0] = clock64() - tstart; a.dbg[1] = w_de; a.dbg[2] = w_af; }
    }
  }
  } else if (warp < 8) {
    // ===================== operand warps: gout pixels -> tf32 hi/lo -> tensor memory =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
    const int q = (int)threadIdx.x - 128;                       // TMEM lane
    const uint32_t t_lane = tmem_a0 + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t it = 0;
    long long w_ae = 0;
    const long long tstart = clock64();
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
        { const long long t0_ = clock64(); ptx::mbar_wait(a_empty(s), ((it >> 1) & 1u) ^ 1u); w_ae += clock64() - t0_; }
        ptx::tc_fence_after();
        const uint32_t ta = t_lane + (uint32_t)s * 2u * (uint32_t)a.Kpad;
        for (int c0 = 0; c0 < a.Kpad; c0 += 16) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; j++) v[j] = (ok && c0 + j < a.CO) ? __ldg(src + (int64_t)(c0 + j) * HWo) : 0.f;
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; j++) {
            const float h = ptx::to_tf32_rna(v[j]);
            const float l = ptx::to_tf32_rna(v[j] - h);
            hi[j] = __float_as_uint(h);
            lo[j] = (fabsf(v[j]) < __int_as_float(0x7f800000)) ? __float_as_uint(l) : 0u;
          }
          ptx::tmem_st_32x16(ta + (uint32_t)c0, hi);
          ptx::tmem_st_32x16(ta + (uint32_t)a.Kpad + (uint32_t)c0, lo);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a_full(s)) : "memory");
      }
    }
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 128) { a.dbg[3] = clock64() - tstart; a.dbg[4] = w_ae; }
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
    const float* zero_p = img_p + img_floats;                   // one word that stays 0.0f
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
    const long long tstart = clock64();
    for (int64_t g = g0; g < a.ngroups; g += gstep) {
      const int64_t n0 = g * a.ipg;
      const int imgs = (a.N - n0 < a.ipg) ? (int)(a.N - n0) : a.ipg;
      const int npix = imgs * HWo;
      const int nout = imgs * HWi;                              // input pixels per channel in this group
      for (int t = 0; t < a.tpg; t++, it++) {
        const int s = it & 1;
        const int q_lo = t * 128, q_hi = (q_lo + 128 < npix) ? q_lo + 128 : npix;      // group pixels this tile holds
        const bool whole = (q_lo == 0 && q_hi == npix);         // the tile holds every pixel of the group: no range test
        { const long long t0_ = clock64(); ptx::mbar_wait(d_full(s), (it >> 1) & 1u); w_df += clock64() - t0_; }
        const long long tp1 = clock64();
        ptx::tc_fence_after();
        const uint32_t td = tmem_base + ((uint32_t)(wq * 32) << 16) + 128u * (uint32_t)s;
        // phase 1: D -> col_s[column][pixel] (lanes = consecutive pixels: conflict-free)
        if (cgrp * 32 < ncols) {
          uint32_t r32[32];                                      // this warp's 32 columns in one go (reads past ncols are harmless)
          ptx::tmem_ld_32x32(td + (uint32_t)(cgrp * 32), r32);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; e++)
            if (cgrp * 32 + e < ncols && ppos >= 0) asm volatile("st.shared.b32 [%0], %1;" ::"r"(col_base + (uint32_t)((cgrp * 32 + e) * a.colL + ppos) * 4u), "r"(r32[e]) : "memory");
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(d_empty(s)) : "memory");   // accumulator free again
        bar_sync_named(1, 512);
        const long long tp2 = clock64();
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
        w_p2 += clock64() - tp2;
      }
      const long long tfl = clock64();
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
      w_fl += clock64() - tfl;
    }
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 256) { a.dbg[5] = clock64() - tstart; a.dbg[6] = w_df; a.dbg[7] = w_p1; a.dbg[8] = w_p2; a.dbg[9] = w_fl; }
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
  static int dbg_env = -1;
  if (dbg_env < 0) { const char* e = getenv("AM_CONVTC_DEBUG"); dbg_env = (e && e[0] == '1') ? 1 : 0; }
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

}  // namespace am
