// float32 conv2d forward / stride-1 data gradient as a fused implicit GEMM on the 5th-gen tensor cores
// (tcgen05 kind::tf32, 3xTF32 split, fp32 accuracy).  The im2col matrix never exists in HBM: the raw NCHW images a
// tile of 128 output pixels needs are bulk-copied (TMA 1-D, cp.async.bulk) into shared memory once, and the gather
// warps expand them from there, 32 k at a time, split into tf32 hi/lo, straight into TENSOR MEMORY (tcgen05.st): the
// MMA reads its A operand from TMEM (lane = pixel, column = k), so the expansion costs no shared-memory bandwidth and
// needs no swizzle; the weights (small, L2-resident) come pre-split via 2-D TMA into their own, deeper ring.
//
//   D[p, co] = sum_k im2col[p, k] * W[co, k],  M = 128 output pixels per tile (TMEM lanes), N = Cout (<= 64 per CTA),
//   K = C*kH*kW.  GEMM view of conv.nim:81-106 (forward) and of col2im(W^T gout) (conv.nim:136-139, gather form).
//
// Persistent, warp-specialised CTA, tiles strided over the grid:
//   warp 0        TMA producer for the weight tiles (hi / lo planes, K-major, packed by conv_tc_pack_weights_kernel)
//   warp 1        UMMA issuer: per 8-wide k step three tcgen05.mma (lo*hi, hi*lo, hi*hi), accumulation chains of
//                 `flush_kb` k blocks into one of two TMEM buffers (the tensor core truncates when it accumulates,
//                 see gemm_f32_tc.cu)
//   warp 2        TMEM allocation
//   warp 3        bulk-copy producer for the raw images of the NEXT tile (double-buffered)
//   warps 4..15   `groups` (<= 3) gather groups of 128 threads (thread r <-> output pixel r <-> TMEM lane r); group g
//                 expands the k blocks g, g + groups, ... so several of the 4 operand stages are being filled at once
//   last 8 warps  (lane quarter x column half) drain finished chains (tcgen05.ld) into register accumulators with round-to-nearest adds, then the
//                 epilogue: + bias, NCHW stores (lanes = consecutive pixels: coalesced)
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


struct ConvTcArgs {
  const float* x;       // [N][C][H][W]        (grad_output for dgrad)
  const float* bias;    // [CO] or null
  float* y;             // [N][CO][HO][WO]     (grad_input for dgrad)
  const int2* tab;      // k -> {offset ci*H*W + kh*dH*W + kw*dW, (kh*dH << 16) | (kw*dW)}; Kpad entries
  int64_t P;            // total output pixels N*HO*WO
  int64_t N;
  int C, H, W, CO, HO, WO, padH, padW, sH, sW;
  int K, kblocks;       // K' = C*kH*kW, ceil(K'/32)
  int NP;               // Cout padded to a multiple of 16 (<= 64): UMMA N and weight-tile rows
  int flush_kb;
  int ntiles;
  int CHW;              // floats per input image
  int stages;           // pipeline depth of the operand ring
  int groups;           // gather groups (of 128 threads)
  uint32_t raw_bytes;   // bytes of one raw-image buffer (max images a tile touches * CHW * 4, rounded up)
  long long* dbg;       // optional: per-role wait-cycle counters of CTA 0 (AM_CONVTC_DEBUG=1)
  int relu;             // fused activation: y = max(0, conv + bias)
  int hi_res;           // the hi weight plane of every k block stays resident in shared memory; only the lo plane is streamed
  int skip;             // experiments (knob convtc_debug >> 1): bit 0 = gather warps skip their loads / stores to TMEM, bit 1 = no output stores
};

constexpr int kCtABytes = 128 * 128;      // one plane of the im2col tile: 128 pixels x 32 floats

__global__ void conv_tc_pack_weights_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo,
                                            int CO, int Rpad, int K, int Kpad, int Cin, int kH, int kW, int64_t w_off,
                                            int64_t w_sco, int64_t w_sci, int64_t w_skh, int64_t w_skw) {
  const int64_t total = (int64_t)Rpad * Kpad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i / Kpad), k = (int)(i - (int64_t)co * Kpad);
    float v = 0.f;
    if (co < CO && k < K) {
      const int ci = k / (kH * kW), r = k - ci * (kH * kW), kh = r / kW, kw = r - kh * kW;
      v = w[w_off + co * w_sco + ci * w_sci + kh * w_skh + kw * w_skw];
    }
    float h = v, l = 0.f;
    if (isfinite(v)) { h = ptx::to_tf32_rna(v); l = ptx::to_tf32_rna(v - h); if (!isfinite(h)) { h = v; l = 0.f; } }
    hi[i] = h;
    lo[i] = l;
  }
}

__global__ void conv_tc_table_kernel(int2* tab, int K, int Kpad, int kH, int kW, int H, int W, int dH, int dW, int checked) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < Kpad; k += gridDim.x * blockDim.x) {
    if (k < K) {
      const int ci = k / (kH * kW), r = k - ci * (kH * kW), kh = r / kW, kw = r - kh * kW;
      tab[k] = make_int2(ci * H * W + kh * dH * W + kw * dW, ((kh * dH) << 16) | (kw * dW));
    } else {
      // padded taps: the weights are zero; the checked variant also reads a zero (h0 + 32767 is never inside the image),
      // the unchecked one re-reads tap 0 of the pixel's own window
      tab[k] = make_int2(0, checked ? 0x7fff7fff : 0);
    }
  }
}

// CHECK = the convolution has padding: taps outside the image read as zero (bounds test per element)
// Tensor-memory plan (512 columns): FOUR accumulator buffers of up to 64 columns + four im2col operand stages of 64 columns
// (hi 32 | lo 32).  An accumulation chain is only 2 k blocks = 24 MMAs (~1000 cycles) long for fp32 accuracy, and draining
// one (barrier wake-up, two dependent tcgen05.ld round trips, 64 adds, arrive) takes about as long: with two buffers the
// tensor pipe waited for the drain of chain c-2 before it could start chain c (measured: 134 us at chains of 2 k blocks,
// 101 us at chains of 8 — at four times the rounding error).  Four buffers let the MMAs run three chains ahead.
constexpr int kCtAStages = 4;
constexpr int kCtAccBufs = 4;
constexpr int kCtAccCols = 64 * kCtAccBufs;
constexpr int kCtMaxBStages = 8;

#define CT_TWAIT(counter, bar, ph) do { const long long t0_ = pclk(); ptx::mbar_wait(bar, ph); counter += pclk() - t0_; } while (0)

template <bool CHECK>
__global__ void __launch_bounds__(768, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo, const ConvTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int SB = a.stages;                                              // weight ring depth
  const uint32_t b_bytes = (uint32_t)a.NP * 128u;                     // one weight plane per stage
  // ring stage: [hi plane | lo plane], or the lo plane alone when the hi planes of all k blocks are resident (hi_res): the
  // weight stream from L2 is what slows the MMAs down (profiles/r02_conv_tc_issue_analysis.md), this halves it
  const uint32_t stage_bytes = a.hi_res ? b_bytes : 2u * b_bytes;
  auto stage_base = [&](int s) { return smem_base + (uint32_t)s * stage_bytes; };
  const uint32_t OFF_BHI = 0, OFF_BLO = a.hi_res ? 0u : b_bytes;
  const uint32_t hi_base = smem_base + (uint32_t)SB * stage_bytes;      // hi_res: [kblocks][NP x 32] resident hi planes
  const uint32_t raw_base = hi_base + (a.hi_res ? (uint32_t)a.kblocks * b_bytes : 0u);     // two raw-image buffers
  const uint32_t bar_base = raw_base + 2u * a.raw_bytes;
  auto full_a = [&](int s) { return bar_base + 8u * s; };               // 4
  auto empty_a = [&](int s) { return bar_base + 8u * (4 + s); };        // 4
  auto full_b = [&](int s) { return bar_base + 8u * (8 + s); };         // 8
  auto empty_b = [&](int s) { return bar_base + 8u * (16 + s); };       // 8
  auto tfull_bar = [&](int b) { return bar_base + 8u * (24 + b); };     // 4
  auto tempty_bar = [&](int b) { return bar_base + 8u * (28 + b); };    // 4
  auto raw_full = [&](int b) { return bar_base + 8u * (32 + b); };
  auto raw_empty = [&](int b) { return bar_base + 8u * (34 + b); };
  const uint32_t tmem_slot = bar_base + 8u * 36;
  const uint32_t hi_full = bar_base + 8u * 37;                          // hi_res: the resident planes have landed
  const uint32_t bias_s = bar_base + 8u * 38;                           // float bias[64] (zero when there is none)
  const uint32_t tab_s = bias_s + 256u;                                 // int2 tab[kblocks*32]

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = ptx::lane_id();
  const int nkb = a.kblocks;
  const int G = a.groups;
  const int acc_warp0 = 16;                               // the last eight of the 24 launched warps
  const uint32_t tmem_cols = 512u;                        // accumulators (256) + 4 operand stages (256)
  const int64_t HW = (int64_t)a.HO * a.WO;

  if (warp == 0 && ptx::elect_one()) { ptx::prefetch_tensormap(&tmWhi); ptx::prefetch_tensormap(&tmWlo); }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < kCtAStages; s++) {
      ptx::mbar_init(full_a(s), 128);      // every thread of the filling group arrives after its tcgen05.st completed
      ptx::mbar_init(empty_a(s), 1);       // tcgen05.commit
    }
    for (int s = 0; s < SB; s++) {
      ptx::mbar_init(full_b(s), 1);        // TMA transaction bytes
      ptx::mbar_init(empty_b(s), 1);       // tcgen05.commit
    }
    for (int b = 0; b < kCtAccBufs; b++) { ptx::mbar_init(tfull_bar(b), 1); ptx::mbar_init(tempty_bar(b), a.NP > 32 ? 8 : 4); }
    for (int b = 0; b < 2; b++) { ptx::mbar_init(raw_full(b), 1); ptx::mbar_init(raw_empty(b), 128u * (uint32_t)G); }
    ptx::mbar_init(hi_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<1>(tmem_slot, tmem_cols);
  if (threadIdx.x < 64) {
    const float bv = (a.bias != nullptr && (int)threadIdx.x < a.CO) ? a.bias[threadIdx.x] : 0.f;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + 4u * threadIdx.x), "f"(bv));
  }
  for (int i = threadIdx.x; i < nkb * 32; i += blockDim.x) {            // k-decode table -> shared memory
    const int2 e = a.tab[i];
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(tab_s + 8u * i), "r"(e.x), "r"(e.y));
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int flush = a.flush_kb;
  const int chains_per_tile = (nkb + flush - 1) / flush;

  // register budget (24 warps launched at 80 per thread = 61440): control warps 56, up to 12 gather warps 72 (idle ones 40),
  // eight accumulate warps 104  (7168 + 27648 + 26624 = 61440)
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ===================== TMA producer: weight tiles (own ring, runs ahead of the operand stages) =====================
      if (ptx::elect_one()) {
        int s = 0; uint32_t ph = 1u;                  // ring position and wait parity, advanced without divisions
        long long w_eb = 0;
        const long long tstart = pclk();
        if (a.hi_res) {
          ptx::mbar_arrive_expect_tx(hi_full, (uint32_t)nkb * b_bytes);
          for (int kb = 0; kb < nkb; kb++) ptx::tma_load_2d(hi_base + (uint32_t)kb * b_bytes, &tmWhi, hi_full, kb * 32, 0);
        }
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
          for (int kb = 0; kb < nkb; kb++) {
            CT_TWAIT(w_eb, empty_b(s), ph);
            const uint32_t sb = stage_base(s);
            ptx::mbar_arrive_expect_tx(full_b(s), stage_bytes);
            if (!a.hi_res) ptx::tma_load_2d(sb + OFF_BHI, &tmWhi, full_b(s), kb * 32, 0);
            ptx::tma_load_2d(sb + OFF_BLO, &tmWlo, full_b(s), kb * 32, 0);
            if (++s == SB) { s = 0; ph ^= 1u; }
          }
        }
        if (a.dbg && blockIdx.x == 0) { a.dbg[0] = pclk() - tstart; a.dbg[1] = w_eb; }
      }
    } else if (warp == 1) {
      // ===================== UMMA issuer: A from tensor memory, B from shared memory =====================
      if (ptx::elect_one()) {
        const uint64_t dhi = ptx::umma_desc_hi(1024, 128);
        const uint32_t idesc = ptx::umma_idesc_tf32(128, (uint32_t)a.NP);
        uint32_t chain = 0;
        // ring positions / wait parities of the two operand rings advance by compare-and-wrap: the issuing thread runs
        // nothing but a dozen uniform-register adds between two batches of MMAs (a division by the run-time ring depth
        // plus the multiply-shift by 6 used to cost ~90 dependent instructions per k block: 66 cycles per MMA where the
        // same issue pattern alone sustains 38, am_microbench 27/31)
        int sa = 0, sbi = 0; uint32_t pa = 0u, pb = 0u;
        uint32_t a_hi0 = tmem_base + (uint32_t)kCtAccCols;
        uint64_t b_hi0 = ptx::umma_desc(dhi, a.hi_res ? hi_base : stage_base(0) + OFF_BHI), b_lo0 = ptx::umma_desc(dhi, stage_base(0) + OFF_BLO);
        const uint64_t b_hi_first = b_hi0, b_lo_first = b_lo0, b_step = (uint64_t)(stage_bytes >> 4), b_step_hi = (uint64_t)(b_bytes >> 4);
        if (a.hi_res) ptx::mbar_wait(hi_full, 0u);
        long long w_te = 0, w_fa = 0, w_fb = 0;
        const long long tstart = pclk();
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
          int kb = 0;
          if (a.hi_res) b_hi0 = b_hi_first;                   // resident planes are indexed by the k block of the tile
          for (int c = 0; c < chains_per_tile; c++, chain++) {
            const int buf = chain & (kCtAccBufs - 1);
            CT_TWAIT(w_te, tempty_bar(buf), ((chain / kCtAccBufs) & 1u) ^ 1u);
            ptx::tc_fence_after();
            const uint32_t d = tmem_base + (uint32_t)buf * 64u;
            const int kb_end = (kb + flush < nkb) ? kb + flush : nkb;
            for (uint32_t acc_on = 0u; kb < kb_end; kb++) {
              CT_TWAIT(w_fa, full_a(sa), pa);
              CT_TWAIT(w_fb, full_b(sbi), pb);
              ptx::tc_fence_after();
              const uint32_t a_lo0 = a_hi0 + 32u;
#pragma unroll
              for (int k8 = 0; k8 < 4; k8++) {
                // the start-address field of a K-major SW128 descriptor counts 16-byte units: +2 per 8-wide k step
                const uint64_t b_hi = b_hi0 + 2u * k8, b_lo = b_lo0 + 2u * k8;
                ptx::umma_tf32_ts(d, a_lo0 + 8u * k8, b_hi, idesc, k8 == 0 ? acc_on : 1u);
                ptx::umma_tf32_ts(d, a_hi0 + 8u * k8, b_lo, idesc, 1u);
                ptx::umma_tf32_ts(d, a_hi0 + 8u * k8, b_hi, idesc, 1u);
              }
              acc_on = 1u;
              ptx::umma_commit<1>(empty_a(sa));
              ptx::umma_commit<1>(empty_b(sbi));
              a_hi0 += 64u;
              if (++sa == kCtAStages) { sa = 0; pa ^= 1u; a_hi0 = tmem_base + (uint32_t)kCtAccCols; }
              b_lo0 += b_step;
              b_hi0 += a.hi_res ? b_step_hi : b_step;
              if (++sbi == SB) { sbi = 0; pb ^= 1u; b_lo0 = b_lo_first; if (!a.hi_res) b_hi0 = b_hi_first; }
            }
            ptx::umma_commit<1>(tfull_bar(buf));
          }
        }
        if (a.dbg && blockIdx.x == 0) { a.dbg[2] = pclk() - tstart; a.dbg[3] = w_te; a.dbg[4] = w_fa; a.dbg[5] = w_fb; }
      }
    } else if (warp == 3) {
      // ===================== raw-image producer: whole images of the tile, one contiguous bulk copy =====================
      if (ptx::elect_one()) {
        uint32_t ti = 0;
        long long w_re = 0;
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ti++) {
          const int b = ti & 1;
          CT_TWAIT(w_re, raw_empty(b), ((ti >> 1) & 1u) ^ 1u);
          const int64_t p0 = (int64_t)tile * 128;
          const int64_t p1 = (p0 + 127 < a.P - 1) ? p0 + 127 : a.P - 1;
          const int64_t n0 = p0 / HW, n1 = p1 / HW;
          const uint32_t bytes = (uint32_t)((n1 - n0 + 1) * a.CHW) * 4u;
          ptx::mbar_arrive_expect_tx(raw_full(b), bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(raw_base + (uint32_t)b * a.raw_bytes), "l"(a.x + n0 * a.CHW), "r"(bytes), "r"(raw_full(b)) : "memory");
        }
        if (a.dbg && blockIdx.x == 0) a.dbg[6] = w_re;
      }
    }
  } else if (warp < 4 + 4 * G) {
    // ===================== gather groups: expand the raw images into the im2col operand (tensor memory) =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int g = (warp - 4) >> 2;
    const int r = (int)(threadIdx.x - 128) & 127;          // pixel row of the tile = TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)kCtAccCols;
    uint32_t ti = 0;
    long long w_rf = 0, w_ea = 0, w_st = 0;
    const long long tstart = pclk();
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ti++) {
      const int64_t p0 = (int64_t)tile * 128;
      const int64_t p = p0 + r;
      const bool pok = p < a.P;
      const int64_t pp = pok ? p : p0;
      const int64_t n0 = p0 / HW;
      const int64_t n = pp / HW;
      const int rem = (int)(pp - n * HW);
      const int ho = rem / a.WO, wo = rem - ho * a.WO;
      const int h0 = (pok || !CHECK) ? ho * a.sH - a.padH : -100000, w0 = wo * a.sW - a.padW;
      const int b = ti & 1;
      // byte address of (image n, channel 0, h0, w0) in the raw buffer (may point before the image when padded)
      const uint32_t xb = raw_base + (uint32_t)b * a.raw_bytes + (uint32_t)(((int)(n - n0) * a.CHW + h0 * a.W + w0) * 4);
      CT_TWAIT(w_rf, raw_full(b), (ti >> 1) & 1u);
      for (int kb = g; kb < nkb; kb += G) {
        const uint32_t it = ti * (uint32_t)nkb + (uint32_t)kb;
        const int sa = it % kCtAStages;
        CT_TWAIT(w_ea, empty_a(sa), ((it / kCtAStages) & 1u) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t ta = t_lane + 64u * (uint32_t)sa;
#pragma unroll
        for (int half = 0; half < 2; half++) {
          if (a.skip & 1) break;
          float v[16];
#pragma unroll
          for (int j2 = 0; j2 < 8; j2++) {                 // two table entries per 128-bit broadcast load
            int e0x, e0y, e1x, e1y;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(e0x), "=r"(e0y), "=r"(e1x), "=r"(e1y)
                         : "r"(tab_s + 8u * (uint32_t)(kb * 32 + half * 16 + 2 * j2)));
            float x0, x1;
            if (CHECK) {
              const int ha = h0 + (e0y >> 16), wa = w0 + (e0y & 0xffff), hb = h0 + (e1y >> 16), wb = w0 + (e1y & 0xffff);
              x0 = 0.f; x1 = 0.f;
              if ((unsigned)ha < (unsigned)a.H && (unsigned)wa < (unsigned)a.W) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x0) : "r"(xb + 4u * (uint32_t)e0x));
              if ((unsigned)hb < (unsigned)a.H && (unsigned)wb < (unsigned)a.W) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x1) : "r"(xb + 4u * (uint32_t)e1x));
            } else {
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x0) : "r"(xb + 4u * (uint32_t)e0x));
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x1) : "r"(xb + 4u * (uint32_t)e1x));
            }
            v[2 * j2] = x0; v[2 * j2 + 1] = x1;
          }
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; j++) {
            ptx::split_tf32(v[j], hi[j], lo[j]);
          }
          ptx::tmem_st_32x16(ta + 16u * half, hi);
          ptx::tmem_st_32x16(ta + 32u + 16u * half, lo);
        }
        { const long long t0_ = pclk(); ptx::tmem_st_wait(); w_st += pclk() - t0_; }
        ptx::tc_fence_before();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_a(sa)) : "memory");
      }
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(raw_empty(b)) : "memory");   // done reading this tile's images
    }
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 128) { a.dbg[7] = pclk() - tstart; a.dbg[8] = w_rf; a.dbg[9] = w_ea; a.dbg[10] = w_st; }
  } else if (warp < acc_warp0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  } else {
    // ===================== accumulate / epilogue warps =====================
    // Eight warps: warp (q, half) owns TMEM lanes 32q..32q+31 (pixels) and accumulator columns 32*half..32*half+31 (output
    // channels).  A chain of 2 k blocks is executed in ~1000 cycles and has to be drained in less: one tcgen05.ld + 32 adds
    // per warp (four warps x 64 columns took two dependent loads and twice the adds: the tensor pipe waited for them)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int q = warp & 3, half = (warp - acc_warp0) >> 2;
    const int r = q * 32 + (int)lane;                      // TMEM lane = pixel row of the tile
    const int c_lo = half * 32;
    if (c_lo < a.NP) {
    uint32_t chain = 0;
    long long w_tf = 0, w_ep = 0;
    const long long tstart = pclk();
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      float acc[32];
#pragma unroll
      for (int i = 0; i < 32; i++) acc[i] = 0.f;
      for (int c = 0; c < chains_per_tile; c++, chain++) {
        const int buf = chain & (kCtAccBufs - 1);
        CT_TWAIT(w_tf, tfull_bar(buf), (chain / kCtAccBufs) & 1u);
        ptx::tc_fence_after();
        uint32_t r0[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 64 + c_lo), r0);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) acc[i] = __fadd_rn(acc[i], __uint_as_float(r0[i]));
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(buf)) : "memory");
      }
      // epilogue: out[n][co][ho][wo] = acc[co] + bias[co]; consecutive lanes = consecutive pixels
      const long long te_ = pclk();
      const int64_t p = (int64_t)tile * 128 + r;
      if (p < a.P && !((a.skip & 2) && acc[0] != 12345.f)) {
        const int64_t n = p / HW, rem = p - n * HW;
        float* yp = a.y + (n * a.CO + c_lo) * HW + rem;
        // the bias comes from shared memory: a global load here could alias the stores and would serialise them
#pragma unroll
        for (int c4 = 0; c4 < 8; c4++) {
          if (c_lo + c4 * 4 < a.CO) {
            float b0, b1, b2, b3;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3) : "r"(bias_s + 4u * (uint32_t)c_lo + 16u * c4));
            const float bb[4] = {b0, b1, b2, b3};
#pragma unroll
            for (int e = 0; e < 4; e++)
              if (c_lo + c4 * 4 + e < a.CO) {
                const float v = __fadd_rn(acc[c4 * 4 + e], bb[e]);
                yp[(c4 * 4 + e) * HW] = (a.relu && v <= 0.f) ? 0.f : v;
              }
          }
        }
      }
      w_ep += pclk() - te_;
    }
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 512) { a.dbg[11] = pclk() - tstart; a.dbg[12] = w_tf; a.dbg[13] = w_ep; }
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<1>(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_TmapEncodeTiled2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct ConvTcView {           // one convolution in "forward form": y = conv(x, W) with W given as a strided 4-D view
  const float* x; const float* w; const float* bias; float* y;
  int64_t N;
  int C, H, W, CO, kH, kW, padH, padW, sH, sW, dH, dW, HO, WO;
  int64_t w_off, w_sco, w_sci, w_skh, w_skw;
  int relu;
};

static int run_conv_tc(cudaStream_t st, const ConvTcView& v, bool* done) {
  *done = false;
  if (!gemm_f32_tc_available()) return AM_OK;
  const int K = v.C * v.kH * v.kW;
  if (v.CO > 64 || K > 4096) return AM_OK;                       // accumulators: 64 registers per pixel; table in smem
  if (v.kH * v.dH >= 32767 || v.kW * v.dW >= 32767) return AM_OK;
  const int64_t CHW = (int64_t)v.C * v.H * v.W;
  // whole input images are staged in shared memory by 16-byte bulk copies
  if (CHW % 4 != 0 || (reinterpret_cast<uintptr_t>(v.x) & 15) != 0) return AM_OK;
  const int64_t P = v.N * (int64_t)v.HO * v.WO;
  if (P <= 0 || P >= (1ll << 37)) return AM_OK;
  const int Kpad = (int)round_up(K, 32), nkb = Kpad / 32;
  const int NP = (int)round_up(v.CO, 16);
  const int Rpad = 64;                                            // weight planes padded to 64 rows
  // shared-memory plan: weight ring + two raw-image buffers + barriers + k table (the im2col operand lives in TMEM)
  const int64_t HWo = (int64_t)v.HO * v.WO;
  const int64_t max_imgs = (128 % HWo == 0) ? 128 / HWo : 127 / HWo + 2;      // images one 128-pixel tile can touch
  const int64_t raw_bytes = round_up(max_imgs * CHW * 4, 128);
  const size_t b_bytes = (size_t)NP * 128;
  size_t stage_bytes = 2 * b_bytes;
  size_t fixed = 1024 + 8 * 38 + 256 + (size_t)Kpad * 8 + 64 + 2 * (size_t)raw_bytes;
  if (fixed + 2 * stage_bytes > 227 * 1024) return AM_OK;
  // resident hi planes (+ a ring of >= 4 lo planes) when they fit: halves the weight stream every tile re-reads from L2
  const size_t hi_bytes = (size_t)nkb * b_bytes;
  const bool hi_res = tuning(kTuneConvTcHiRes) != 0 && fixed + hi_bytes + 4 * b_bytes <= 227 * 1024;
  if (hi_res) { stage_bytes = b_bytes; fixed += hi_bytes; }
  int stages = (int)((227 * 1024 - fixed) / stage_bytes);
  if (stages > kCtMaxBStages) stages = kCtMaxBStages;
  const bool checked = v.padH != 0 || v.padW != 0;
  // workspace: weight hi/lo planes + k table
  void* ws = nullptr;
  const size_t plane = (size_t)Rpad * Kpad * sizeof(float);
  int rc = workspace(kWsConvW, 2 * plane + (size_t)Kpad * sizeof(int2) + 256, &ws);
  if (rc) return rc;
  float* whi = (float*)ws; float* wlo = whi + (size_t)Rpad * Kpad;
  int2* tab = (int2*)((char*)ws + 2 * plane);
  conv_tc_pack_weights_kernel<<<(unsigned)ceil_div((int64_t)Rpad * Kpad, 256), 256, 0, st>>>(
      v.w, whi, wlo, v.CO, Rpad, K, Kpad, v.C, v.kH, v.kW, v.w_off, v.w_sco, v.w_sci, v.w_skh, v.w_skw);

  conv_tc_table_kernel<<<(unsigned)ceil_div(Kpad, 256), 256, 0, st>>>(tab, K, Kpad, v.kH, v.kW, v.H, v.W, v.dH, v.dW, checked ? 1 : 0);
  g_launch_count += 2;
  AM_CUDA_TRY(cudaGetLastError());

  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  AM_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) return AM_OK;
  auto encode = (PFN_TmapEncodeTiled2)fn;
  CUtensorMap tms[2];
  float* planes[2] = {whi, wlo};
  for (int i = 0; i < 2; i++) {
    cuuint64_t gdim[2] = {(cuuint64_t)Kpad, (cuuint64_t)Rpad};
    cuuint64_t gstride[1] = {(cuuint64_t)Kpad * sizeof(float)};
    cuuint32_t box[2] = {32u, (cuuint32_t)NP};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = encode(&tms[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, planes[i], gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return AM_ERR_CUDA; }
  }
  ConvTcArgs a{};
  a.x = v.x; a.bias = v.bias; a.y = v.y; a.tab = tab; a.P = P; a.N = v.N; a.relu = v.relu;
  a.CHW = (int)CHW; a.stages = stages; a.raw_bytes = (uint32_t)raw_bytes; a.hi_res = hi_res ? 1 : 0;
  const int groups_env = (tuning(kTuneConvTcGroups) >= 1 && tuning(kTuneConvTcGroups) <= 3) ? tuning(kTuneConvTcGroups) : 3;
  a.groups = groups_env;
  const int dbg_env = tuning(kTuneConvTcDebug) & 1;
  a.skip = tuning(kTuneConvTcDebug) >> 1;
  a.dbg = nullptr;
  if (dbg_env) {
    void* base = nullptr;
    if ((rc = workspace(kWsMisc, 64 * sizeof(int) + 1024, &base))) return rc;
    a.dbg = (long long*)((char*)base + 512);
    AM_CUDA_TRY(cudaMemsetAsync(a.dbg, 0, 16 * sizeof(long long), st));
  }
  a.C = v.C; a.H = v.H; a.W = v.W; a.CO = v.CO; a.HO = v.HO; a.WO = v.WO; a.padH = v.padH; a.padW = v.padW; a.sH = v.sH; a.sW = v.sW;
  a.K = K; a.kblocks = nkb; a.NP = NP;
  // chain length: the tensor core truncates when it accumulates, so a chain's rounding error grows with its length, and every
  // chain costs a hand-off to the accumulate warps.  K = C*kH*kW is at most a few thousand here: chains of 4 k blocks (128 k)
  // measure 7e-7 rel. Frobenius on cv2 — the accuracy of the GEMM path at K = 16384 — and 12 % faster than chains of 2 (4e-7).
  const int flush_env = tuning(kTuneConvTcFlushKb) > 0 ? tuning(kTuneConvTcFlushKb) : 4;
  a.flush_kb = flush_env;
  a.ntiles = (int)ceil_div(P, 128);
  const size_t smem = fixed + (size_t)stages * stage_bytes;
  const int grid = a.ntiles < sm_count() ? a.ntiles : sm_count();
  const int nthreads = 768;
  if (checked) {
    AM_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_tc_kernel<true><<<grid, nthreads, smem, st>>>(tms[0], tms[1], a);
  } else {
    AM_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_tc_kernel<false><<<grid, nthreads, smem, st>>>(tms[0], tms[1], a);
  }
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  if (a.dbg) {
    long long h[16];
    AM_CUDA_TRY(cudaMemcpyAsync(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    AM_CUDA_TRY(cudaStreamSynchronize(st));
    fprintf(stderr, "[conv_tc dbg] tiles/cta=%d nkb=%d SB=%d | tma: total %lld wait_empty_b %lld | mma: total %lld wait_tempty %lld wait_full_a %lld wait_full_b %lld | raw: wait_raw_empty %lld | gather0: total %lld wait_raw_full %lld wait_empty_a %lld wait_st %lld | acc: total %lld wait_tfull %lld epilogue %lld\n",
            (a.ntiles + grid - 1) / grid, a.kblocks, a.stages, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10], h[11], h[12], h[13]);
  }
  *done = true;
  return AM_OK;
}

int conv2d_forward_tc_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* input,
                          const float* kernel, const float* bias, float* output, int act, bool* done) {
  ConvTcView v{};
  v.relu = act;
  v.x = input; v.w = kernel; v.bias = bias; v.y = output; v.N = d.N;
  v.C = (int)d.C; v.H = (int)d.H; v.W = (int)d.W; v.CO = (int)d.Cout; v.kH = (int)d.kH; v.kW = (int)d.kW;
  v.padH = (int)d.padH; v.padW = (int)d.padW; v.sH = (int)d.strideH; v.sW = (int)d.strideW; v.dH = (int)d.dilH; v.dW = (int)d.dilW;
  v.HO = (int)Ho; v.WO = (int)Wo;
  v.w_off = 0; v.w_sco = d.C * d.kH * d.kW; v.w_sci = d.kH * d.kW; v.w_skh = d.kW; v.w_skw = 1;
  return run_conv_tc(st, v, done);
}

int conv2d_dgrad_tc_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* grad_output,
                        const float* kernel, float* grad_input, bool* done) {
  *done = false;
  if (d.strideH != 1 || d.strideW != 1) return AM_OK;
  ConvTcView v{};
  v.x = grad_output; v.w = kernel; v.bias = nullptr; v.y = grad_input; v.N = d.N;
  v.C = (int)d.Cout; v.H = (int)Ho; v.W = (int)Wo; v.CO = (int)d.C; v.kH = (int)d.kH; v.kW = (int)d.kW;
  v.padH = (int)(d.dilH * (d.kH - 1) - d.padH); v.padW = (int)(d.dilW * (d.kW - 1) - d.padW);
  v.sH = 1; v.sW = 1; v.dH = (int)d.dilH; v.dW = (int)d.dilW; v.HO = (int)d.H; v.WO = (int)d.W;
  v.w_sco = d.kH * d.kW; v.w_sci = d.C * d.kH * d.kW; v.w_skh = -d.kW; v.w_skw = -1;
  v.w_off = (d.kH - 1) * d.kW + (d.kW - 1);
  return run_conv_tc(st, v, done);
}

}  // namespace am
