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
// Persistent, warp-specialised CTA (384 threads).  A CTA owns one CHUNK of input channels (cpc channels = up to 128
// rows r; its W^T slice stays resident in shared memory) and walks over groups of whole images:
//   warp 0      one TMA load of the CTA's W^T chunk (hi / lo planes, K-major, 128B swizzle)
//   warp 1      UMMA issuer: A from tensor memory, B from shared memory, (Kpad/8)*3 tcgen05.mma per tile
//   warp 2      TMEM allocation (2 accumulators of 128 columns + 2 operand stages of 2*Kpad columns)
//   warps 4-7   operand warps: thread q loads gout[n, :, q] (lanes = consecutive pixels: coalesced), splits into
//               tf32 hi / lo and writes its TMEM lane with tcgen05.st
//   warps 8-11  col2im warps: tcgen05.ld the tile's D row, park it in shared memory (col_s[r][q]; the accumulator is
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

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(384, 1)
conv_dgrad_tc_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo, const DgradTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nkb = a.Kpad >> 5;
  const int KK = a.kH * a.kW;
  const int HWi = a.H * a.W, HWo = a.HO * a.WO;
  const uint32_t plane_bytes = (uint32_t)nkb * 16384u;                   // 128 rows x Kpad floats, k blocks of 32
  const uint32_t OFF_BHI = 0, OFF_BLO = plane_bytes;
  const uint32_t col_base = smem_base + 2u * plane_bytes;                 // col_s[cpc*KK <= 128][128]
  const uint32_t col_bytes = (uint32_t)(a.cpc * KK) * 512u;
  const uint32_t img_base = col_base + col_bytes;                         // gin_s[cpc][ipg][H*W]
  const uint32_t img_floats = (uint32_t)(a.cpc * a.ipg * HWi);
  const uint32_t bar_base = (img_base + img_floats * 4u + 15u) & ~15u;
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
      ptx::mbar_init(d_full(s), 1); ptx::mbar_init(d_empty(s), 4);
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
      for (int64_t g = g0; g < a.ngroups; g += gstep) {
        for (int t = 0; t < a.tpg; t++, it++) {
          const int s = it & 1;
          const uint32_t ph = (it >> 1) & 1u;
          ptx::mbar_wait(d_empty(s), ph ^ 1u);
          ptx::mbar_wait(a_full(s), ph);
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
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== operand warps: gout pixels -> tf32 hi/lo -> tensor memory =====================
    const int q = (int)threadIdx.x - 128;                       // TMEM lane
    const uint32_t t_lane = tmem_a0 + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t it = 0;
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
        ptx::mbar_wait(a_empty(s), ((it >> 1) & 1u) ^ 1u);
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
  } else if (warp >= 8) {
    // ===================== col2im warps =====================
    const int tid = (int)threadIdx.x - 256;                     // 0..127 = TMEM lane = pixel of the tile
    const int wq = warp & 3;
    const bool unit = a.sH == 1 && a.sW == 1 && a.dH == 1 && a.dW == 1;
    const int ncols = cpc_here * KK;                            // D columns in use
    uint32_t it = 0;
    for (int64_t g = g0; g < a.ngroups; g += gstep) {
      const int64_t n0 = g * a.ipg;
      const int imgs = (a.N - n0 < a.ipg) ? (int)(a.N - n0) : a.ipg;
      const int npix = imgs * HWo;
      const int nout = imgs * HWi;                              // input pixels per channel in this group
      for (int t = 0; t < a.tpg; t++, it++) {
        const int s = it & 1;
        const int q_lo = t * 128, q_hi = (q_lo + 128 < npix) ? q_lo + 128 : npix;      // group pixels this tile holds
        const bool whole = (q_lo == 0 && q_hi == npix);         // the tile holds every pixel of the group: no range test
        ptx::mbar_wait(d_full(s), (it >> 1) & 1u);
        ptx::tc_fence_after();
        const uint32_t td = tmem_base + ((uint32_t)(wq * 32) << 16) + 128u * (uint32_t)s;
        // phase 1: this thread's pixel row of D -> col_s[column][pixel] (lanes = consecutive pixels: conflict-free)
        for (int j0 = 0; j0 < ncols; j0 += 8) {
          uint32_t r8[8];
          tmem_ld_32x8(td + (uint32_t)j0, r8);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; e++)
            if (j0 + e < ncols) asm volatile("st.shared.b32 [%0], %1;" ::"r"(col_base + (uint32_t)((j0 + e) * 128 + tid) * 4u), "r"(r8[e]) : "memory");
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(d_empty(s)) : "memory");   // accumulator free again
        bar_sync_named(1, 128);
        // phase 2: every thread owns the input pixels o = tid, tid + 128, ... : sum the taps that land there, in a
        // fixed order (kh, kw ascending; tiles ascending) -> deterministic
        for (int o = tid; o < nout; o += 128) {
          const int il = o / HWi, rem = o - il * HWi, h = rem / a.W, w = rem - h * a.W;
          if (unit) {
            int kh_lo = h + a.padH - (a.HO - 1); if (kh_lo < 0) kh_lo = 0;
            int kh_hi = h + a.padH; if (kh_hi > a.kH - 1) kh_hi = a.kH - 1;
            int kw_lo = w + a.padW - (a.WO - 1); if (kw_lo < 0) kw_lo = 0;
            int kw_hi = w + a.padW; if (kw_hi > a.kW - 1) kw_hi = a.kW - 1;
            const int base0 = il * HWo + (h + a.padH) * a.WO + (w + a.padW) - q_lo;   // pixel index of tap (0,0) inside the tile
            for (int cl = 0; cl < cpc_here; cl++) {
              float sum = 0.f;
              for (int kh = kh_lo; kh <= kh_hi; kh++) {
                int px = base0 - kh * a.WO - kw_lo;
                uint32_t ad = col_base + (uint32_t)(((cl * KK + kh * a.kW + kw_lo) * 128) + px) * 4u;
                for (int kw = kw_lo; kw <= kw_hi; kw++, px--, ad += 127u * 4u) {
                  if (whole || (px >= 0 && px < q_hi - q_lo)) {
                    float x;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(ad));
                    sum = __fadd_rn(sum, x);
                  }
                }
              }
              const uint32_t ga = img_base + (uint32_t)(cl * a.ipg * HWi + o) * 4u;
              if (t != 0) {
                float old;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(old) : "r"(ga));
                sum = __fadd_rn(old, sum);
              }
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(ga), "f"(sum) : "memory");
            }
          } else {
            for (int cl = 0; cl < cpc_here; cl++) {
              float sum = 0.f;
              for (int kh = 0; kh < a.kH; kh++) {
                const int hh = h + a.padH - kh * a.dH;
                if (hh < 0) break;
                if (hh % a.sH) continue;
                const int ho = hh / a.sH;
                if (ho >= a.HO) continue;
                for (int kw = 0; kw < a.kW; kw++) {
                  const int ww = w + a.padW - kw * a.dW;
                  if (ww < 0) break;
                  if (ww % a.sW) continue;
                  const int wo = ww / a.sW;
                  if (wo >= a.WO) continue;
                  const int px = il * HWo + ho * a.WO + wo - q_lo;
                  if (px >= 0 && px < q_hi - q_lo) {
                    float x;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(col_base + (uint32_t)((cl * KK + kh * a.kW + kw) * 128 + px) * 4u));
                    sum = __fadd_rn(sum, x);
                  }
                }
              }
              const uint32_t ga = img_base + (uint32_t)(cl * a.ipg * HWi + o) * 4u;
              if (t != 0) {
                float old;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(old) : "r"(ga));
                sum = __fadd_rn(old, sum);
              }
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(ga), "f"(sum) : "memory");
            }
          }
        }
        bar_sync_named(1, 128);                                 // col_s may be overwritten by the next tile
      }
      // flush the finished images (each thread wrote exactly the elements it reads back: no barrier needed)
      for (int cl = 0; cl < cpc_here; cl++) {
        for (int o = tid; o < nout; o += 128) {
          const int il = o / HWi, rem = o - il * HWi;
          float x;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(img_base + (uint32_t)(cl * a.ipg * HWi + o) * 4u));
          a.gin[((n0 + il) * a.C + ci0 + cl) * (int64_t)HWi + rem] = x;
        }
      }
    }
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
                               const float* kernel, float* grad_input, bool* done) {
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
  const int sms = sm_count();
  if (a.nchunks > sms) return AM_OK;
  const size_t plane = (size_t)(a.Kpad / 32) * 16384;
  const size_t smem = 1024 + 2 * plane + (size_t)a.cpc * KK * 512 + (size_t)a.cpc * a.ipg * HWi * 4 + 16 + 8 * 10 + 64;
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
  int grid = (sms / a.nchunks) * a.nchunks;
  if ((int64_t)grid > a.ngroups * a.nchunks) grid = (int)(a.ngroups * a.nchunks);
  AM_CUDA_TRY(cudaFuncSetAttribute(conv_dgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv_dgrad_tc_kernel<<<grid, 384, smem, st>>>(tms[0], tms[1], a);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  *done = true;
  return AM_OK;
}

}  // namespace am
