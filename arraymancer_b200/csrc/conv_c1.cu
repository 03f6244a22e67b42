// Single-input-channel float32 convolution (LeNet cv1 class: C = 1, stride 1, dilation 1, 3x3 / 5x5 taps) — the shape
// where the im2col GEMM of the reference (nn_primitives/fallback/conv.nim:81-140) is K = kH*kW = 25 deep and Cout = 20
// wide: far too thin for the tensor cores, and 12 flop per byte, i.e. on the HBM / FFMA ridge.  Two kernels, each
// reading every tensor ONCE, staged by TMA:
//
//   conv_c1_forward_kernel   one cp.async.bulk.tensor (3-D box over {W, H, image}) per group of images brings the raw
//                            images into shared memory with the zero padding produced by the out-of-bounds fill of the
//                            box (start coordinate -pad); two buffers, full / empty mbarriers, no __syncthreads in the
//                            loop.  A thread owns 4 consecutive output pixels of one row, holds its kH x (4 + kW - 1)
//                            input window in registers and loops over the output channels with the kH*kW weights of a
//                            channel coming from shared memory as broadcast 128-bit loads: 100 FMAs per 7 LDS + one
//                            128-bit store of bias (+ ReLU) fused results (71 % of the executed instructions are FFMA).
//   conv_c1_backward_kernel  data gradient, weight gradient and bias gradient of an image from ONE staged copy of its
//                            grad_output (the reference makes three passes: two GEMMs + col2im + a reduction,
//                            conv.nim:129-140, nnp_convolution.nim:91-94).  The image is consumed in two half-steps of
//                            ceil(Cout/2) channel planes: a 4-D TMA box {WO + pads, HO + skew rows, planes, 1} lands
//                            each half column-padded (OOB fill again) while the other half is being used.  Warp-
//                            specialised consumers: the first warps own 4 input pixels each and gather the taps of
//                            every plane (gather form of col2im: no atomics), tap row outermost so the row test runs
//                            kH times per image; the other warps own one (plane, kh, row quarter) and slide the kW taps
//                            over the input row, accumulating in registers ACROSS the images of the persistent CTA.
//                            The four row quarters are added in shared memory in fixed order at the end, the per-CTA
//                            partials by c1_reduce_kernel (fixed order: deterministic).
#include <cuda.h>
#include <type_traits>

#include "am_common.cuh"
#include "gemm_dispatch.h"
#include "ptx_sm100.cuh"

namespace am {

int make_tmap_f32_nd(CUtensorMap* tm, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box);     // gemm_f32_tc.cu

struct C1Args {
  const float* w;      // [CO][1][KH][KW]
  const float* bias;   // [CO] or null
  float* y;            // forward: [N][CO][HO][WO]
  float* gi;           // backward: [N][1][H][W] or null
  float* part;         // backward: [CTAs][CO][KH*KW + 1] partial weight / bias gradients, or null
  int64_t N;
  int H, W, CO, padH, padW, HO, WO;
  int XP;              // row pitch (floats, multiple of 4) of a staged zero-padded input image
  int XROWS;           // rows of a staged padded image
  int GX;              // 4-pixel groups per output row (fwd) / per input row (bwd dgrad)
  int IMGS;            // forward: images per CTA iteration
  int GP, GROWS;       // backward: row pitch (floats) and rows of a staged column-padded grad_output plane
  int want_gi, want_gw;
};

// ------------------------------------------------------------------ forward
template <int KH, int KW, bool RELU>
__global__ void __launch_bounds__(320, 3)
conv_c1_forward_kernel(const __grid_constant__ CUtensorMap tmX, const C1Args a) {
  constexpr int KK = KH * KW, KKP = (KK + 3) / 4 * 4, WIN = (4 + KW - 1 + 3) / 4 * 4;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t base_u32 = (ptx::smem_u32(smem_raw) + 127u) & ~127u;
  float* sm = reinterpret_cast<float*>(smem_raw + (base_u32 - ptx::smem_u32(smem_raw)));
  const int img_floats = a.XROWS * a.XP, buf_floats = (a.IMGS * img_floats + 31) & ~31;
  float* xs = sm;                                  // [2][IMGS][XROWS][XP]   (TMA destinations, 128-byte aligned)
  float* ws = xs + 2 * buf_floats;                 // [CO][KKP]
  float* bs = ws + a.CO * KKP;                     // [round_up(CO,4)]
  const uint32_t bar0 = base_u32 + (uint32_t)((2 * buf_floats + a.CO * KKP + ((a.CO + 3) & ~3)) * 4);
  auto full_bar = [&](int b) { return bar0 + 8u * b; };
  auto empty_bar = [&](int b) { return bar0 + 16u + 8u * b; };
  const int tid = threadIdx.x, nt = blockDim.x, nwarps = nt >> 5;
  for (int i = tid; i < a.CO * KKP; i += nt) { const int co = i / KKP, k = i - co * KKP; ws[i] = k < KK ? a.w[co * KK + k] : 0.f; }
  for (int i = tid; i < a.CO; i += nt) bs[i] = a.bias ? a.bias[i] : 0.f;
  if (tid == 0) {
    ptx::prefetch_tensormap(&tmX);
    for (int b = 0; b < 2; b++) { ptx::mbar_init(full_bar(b), 1); ptx::mbar_init(empty_bar(b), nwarps); }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int tasks = a.HO * a.GX;                   // per image
  const int im = tid / tasks, t = tid - im * tasks;
  const int ho = t / a.GX, gx = t - ho * a.GX;
  const bool vec_out = (a.WO % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0);
  const int64_t ngroups = (a.N + a.IMGS - 1) / a.IMGS;
  const int64_t nloc = (ngroups > (int64_t)blockIdx.x) ? (ngroups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t plane = (int64_t)a.HO * a.WO;
  const uint32_t stage_bytes = (uint32_t)(a.IMGS * img_floats * 4);
  auto issue = [&](int64_t k) {                    // thread 0 only: group k of this CTA into buffer k & 1
    const int b = (int)(k & 1);
    const int64_t n0 = (blockIdx.x + k * (int64_t)gridDim.x) * a.IMGS;
    ptx::mbar_arrive_expect_tx(full_bar(b), stage_bytes);
    ptx::tma_load_3d(base_u32 + (uint32_t)(b * buf_floats * 4), &tmX, full_bar(b), -a.padW, -a.padH, (int)n0);
  };
  if (tid == 0 && nloc > 0) issue(0);
  for (int64_t k = 0; k < nloc; k++) {
    const int b = (int)(k & 1);
    ptx::mbar_wait(full_bar(b), (uint32_t)((k >> 1) & 1));
    if (tid == 0 && k + 1 < nloc) {
      if (k + 1 >= 2) ptx::mbar_wait(empty_bar(b ^ 1), (uint32_t)((((k + 1) >> 1) - 1) & 1));   // consumers left that buffer
      issue(k + 1);
    }
    const int64_t n0 = (blockIdx.x + k * (int64_t)gridDim.x) * a.IMGS;
    const int imgs = (int)((a.N - n0 < a.IMGS) ? a.N - n0 : a.IMGS);
    if (im < imgs) {
      // the thread's input window: KH rows x WIN columns starting at padded column 4*gx
      float win[KH][WIN];
      const float* xw = xs + b * buf_floats + im * img_floats + ho * a.XP + 4 * gx;
#pragma unroll
      for (int r = 0; r < KH; r++)
#pragma unroll
        for (int v = 0; v < WIN / 4; v++) {
          const float4 q = *reinterpret_cast<const float4*>(xw + r * a.XP + 4 * v);
          win[r][4 * v] = q.x; win[r][4 * v + 1] = q.y; win[r][4 * v + 2] = q.z; win[r][4 * v + 3] = q.w;
        }
      float* dst = a.y + ((n0 + im) * a.CO * (int64_t)a.HO + ho) * a.WO + 4 * gx;
      const int nvalid = a.WO - 4 * gx;              // >= 1
      const float* wp = ws;
      for (int co = 0; co < a.CO; co++, dst += plane, wp += KKP) {
        float wv[KKP];
#pragma unroll
        for (int v = 0; v < KKP / 4; v++) {
          const float4 q = *reinterpret_cast<const float4*>(wp + 4 * v);
          wv[4 * v] = q.x; wv[4 * v + 1] = q.y; wv[4 * v + 2] = q.z; wv[4 * v + 3] = q.w;
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < KH; r++)
#pragma unroll
          for (int c = 0; c < KW; c++)
#pragma unroll
            for (int p = 0; p < 4; p++) acc[p] = fmaf(wv[r * KW + c], win[r][p + c], acc[p]);
        const float bv = bs[co];
#pragma unroll
        for (int p = 0; p < 4; p++) {
          acc[p] = __fadd_rn(acc[p], bv);
          if (RELU) acc[p] = (acc[p] <= 0.f) ? 0.f : acc[p];          // value <= 0 -> 0, NaN stays NaN
        }
        if (vec_out) *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else {
#pragma unroll
          for (int p = 0; p < 4; p++) if (p < nvalid) dst[p] = acc[p];
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) ptx::mbar_arrive(empty_bar(b));             // this warp is done reading buffer b
  }
}

// ------------------------------------------------------------------ backward (dgrad + wgrad + bias grad, one pass over grad_output)
// smem: xs [2][XROWS][XP] (zero-padded inputs of two consecutive images) | gs [2][CH][GROWS][GP]: grad_output row r of a
//       plane at r*GP + LPAD, zero columns on both sides and GROWS - HO zero rows (bank skew between planes) |
//       wd [KH][CO][KWP] (weights, tap-row major) | red [RQ][CH*2][KH][KW+1] (final in-CTA reduction) | barriers.
template <int KH, int KW>
__global__ void __launch_bounds__(448, 2)
conv_c1_backward_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const C1Args a,
                        int dgrad_threads, int CH) {
  constexpr int KK = KH * KW, KWP = (KW + 3) / 4 * 4, LPAD = (KW - 1 + 3) / 4 * 4, RQ = 4;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t base_u32 = (ptx::smem_u32(smem_raw) + 127u) & ~127u;
  float* sm = reinterpret_cast<float*>(smem_raw + (base_u32 - ptx::smem_u32(smem_raw)));
  const int x_floats = (a.XROWS * a.XP + 31) & ~31, gplane = a.GROWS * a.GP, g_floats = (CH * gplane + 31) & ~31;
  float* xs = sm;                                           // [2][x_floats]
  float* gs = xs + 2 * x_floats;                            // [2][g_floats]
  float* wd = gs + 2 * g_floats;                            // [KH][CO][KWP]
  float* red = wd + KH * a.CO * KWP;                        // [RQ][2*CH][KH][KW + 1]
  const int red_floats = RQ * 2 * CH * KH * (KW + 1);
  const uint32_t bar0 = base_u32 + (uint32_t)((2 * x_floats + 2 * g_floats + KH * a.CO * KWP + red_floats) * 4);
  auto full_bar = [&](int b) { return bar0 + 8u * b; };
  auto empty_bar = [&](int b) { return bar0 + 16u + 8u * b; };
  const int tid = threadIdx.x, nt = blockDim.x, nwarps = nt >> 5;
  for (int i = tid; i < KH * a.CO * KWP; i += nt) {
    const int kh = i / (a.CO * KWP), rem = i - kh * (a.CO * KWP), co = rem / KWP, kw = rem - co * KWP;
    wd[i] = kw < KW ? a.w[co * KK + kh * KW + kw] : 0.f;
  }
  if (tid == 0) {
    ptx::prefetch_tensormap(&tmX); ptx::prefetch_tensormap(&tmG);
    for (int b = 0; b < 2; b++) { ptx::mbar_init(full_bar(b), 1); ptx::mbar_init(empty_bar(b), nwarps); }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const bool vec_gi = (a.W % 4 == 0) && a.gi && ((reinterpret_cast<uintptr_t>(a.gi) & 15) == 0);
  // ---- roles
  const int dh = tid / a.GX, dgx = tid - dh * a.GX;
  const bool d_on = tid < dgrad_threads && a.want_gi && dh < a.H;
  const int wt = tid - dgrad_threads;
  const int wrq = wt / (CH * KH), wrem = wt - wrq * (CH * KH);
  const int wcl = wrem / KH, wkh = wrem - wcl * KH;         // channel inside the half, tap row
  const bool w_on = wt >= 0 && wrq < RQ && a.want_gw;
  const int rows_q = (a.HO + RQ - 1) / RQ;
  const int wr_beg = wrq * rows_q, wr_end = (wr_beg + rows_q < a.HO) ? wr_beg + rows_q : a.HO;
  float wacc[2][KW];
  float bacc[2] = {0.f, 0.f};
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int c = 0; c < KW; c++) wacc[h][c] = 0.f;
  // data-gradient window geometry: staged column of tap kw for pixel p of the group is c0 + p + (KW-1) - kw
  const int c0 = LPAD + 4 * dgx + a.padW - (KW - 1);
  const int dbase = c0 & ~3, dsh = c0 - dbase;
  float dacc[4] = {0.f, 0.f, 0.f, 0.f};

  const int64_t nloc = (a.N > (int64_t)blockIdx.x) ? (a.N - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // images of this CTA
  const int64_t nsteps = 2 * nloc;
  const uint32_t g_bytes = (uint32_t)(CH * gplane * 4), x_bytes = (uint32_t)(a.XROWS * a.XP * 4);
  auto issue = [&](int64_t s) {                             // thread 0 only
    const int64_t n = blockIdx.x + (s >> 1) * (int64_t)gridDim.x;
    const int half = (int)(s & 1);
    const bool with_x = half == 0 && a.want_gw;
    ptx::mbar_arrive_expect_tx(full_bar(half), g_bytes + (with_x ? x_bytes : 0u));
    if (with_x)
      ptx::tma_load_3d(base_u32 + (uint32_t)((int)((s >> 1) & 1) * x_floats * 4), &tmX, full_bar(half), -a.padW, -a.padH, (int)n);
    ptx::tma_load_4d(base_u32 + (uint32_t)((2 * x_floats + half * g_floats) * 4), &tmG, full_bar(half), -LPAD, 0, half * CH, (int)n);
  };
  if (tid == 0 && nsteps > 0) issue(0);
  for (int64_t s = 0; s < nsteps; s++) {
    const int half = (int)(s & 1);
    ptx::mbar_wait(full_bar(half), (uint32_t)((s >> 1) & 1));
    if (tid == 0 && s + 1 < nsteps) {
      if (s + 1 >= 2) ptx::mbar_wait(empty_bar(half ^ 1), (uint32_t)((((s + 1) >> 1) - 1) & 1));
      issue(s + 1);
    }
    const int co_beg = half * CH, nch = (a.CO - co_beg < CH) ? a.CO - co_beg : CH;
    const float* gbuf = gs + half * g_floats;

    if (d_on) {
      auto body = [&](auto shc) {
        constexpr int SH = decltype(shc)::value;
        constexpr int NV = (SH + 3 + KW - 1) / 4 + 1;             // 16-byte vectors covering window indices 0 .. SH+3+KW-1
#pragma unroll
        for (int kh = 0; kh < KH; kh++) {
          const int r = dh + a.padH - kh;
          if ((unsigned)r < (unsigned)a.HO) {
            const float* gp = gbuf + r * a.GP + dbase;
            const float* wp = wd + (kh * a.CO + co_beg) * KWP;
            for (int cl = 0; cl < nch; cl++, gp += gplane, wp += KWP) {
              float wv[KWP], win[4 * NV];
#pragma unroll
              for (int v = 0; v < KWP / 4; v++) {
                const float4 q = *reinterpret_cast<const float4*>(wp + 4 * v);
                wv[4 * v] = q.x; wv[4 * v + 1] = q.y; wv[4 * v + 2] = q.z; wv[4 * v + 3] = q.w;
              }
#pragma unroll
              for (int v = 0; v < NV; v++) {
                const float4 q = *reinterpret_cast<const float4*>(gp + 4 * v);
                win[4 * v] = q.x; win[4 * v + 1] = q.y; win[4 * v + 2] = q.z; win[4 * v + 3] = q.w;
              }
#pragma unroll
              for (int kw = 0; kw < KW; kw++)
#pragma unroll
                for (int p = 0; p < 4; p++) dacc[p] = fmaf(wv[kw], win[SH + p + KW - 1 - kw], dacc[p]);
            }
          }
        }
      };
      switch (dsh) {
        case 0: body(std::integral_constant<int, 0>{}); break;
        case 1: body(std::integral_constant<int, 1>{}); break;
        case 2: body(std::integral_constant<int, 2>{}); break;
        default: body(std::integral_constant<int, 3>{}); break;
      }
      if (half == 1) {
        const int64_t n = blockIdx.x + (s >> 1) * (int64_t)gridDim.x;
        float* dst = a.gi + n * (int64_t)a.H * a.W + dh * a.W + 4 * dgx;
        if (vec_gi) *reinterpret_cast<float4*>(dst) = make_float4(dacc[0], dacc[1], dacc[2], dacc[3]);
        else {
#pragma unroll
          for (int p = 0; p < 4; p++) if (4 * dgx + p < a.W) dst[p] = dacc[p];
        }
        dacc[0] = dacc[1] = dacc[2] = dacc[3] = 0.f;
      }
    }

    if (w_on && wcl < nch) {
      // gW[co][kh][kw] += sum_{ho,wo} gout[co][ho][wo] * xpad[ho + kh][wo + kw]
      const float* xbuf = xs + (int)((s >> 1) & 1) * x_floats;
      const float* grow = gbuf + wcl * gplane + wr_beg * a.GP + LPAD;
      const float* xrow = xbuf + (wr_beg + wkh) * a.XP;
      float acc[KW], bsum = 0.f;
#pragma unroll
      for (int c = 0; c < KW; c++) acc[c] = 0.f;
      for (int r = wr_beg; r < wr_end; r++, grow += a.GP, xrow += a.XP) {
        float4 xa = *reinterpret_cast<const float4*>(xrow);
        for (int c = 0; c < a.WO; c += 4) {                          // columns beyond WO are zero in gs
          const float4 g4 = *reinterpret_cast<const float4*>(grow + c);
          const float4 xb = *reinterpret_cast<const float4*>(xrow + c + 4);   // XP leaves room for this read
          const float xw[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
          const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
          for (int kw = 0; kw < KW; kw++)
#pragma unroll
            for (int p = 0; p < 4; p++) acc[kw] = fmaf(gv[p], xw[p + kw], acc[kw]);
          if (wkh == 0) bsum += (gv[0] + gv[1]) + (gv[2] + gv[3]);
          xa = xb;
        }
      }
      if (half == 0) {
#pragma unroll
        for (int c = 0; c < KW; c++) wacc[0][c] += acc[c];
        bacc[0] += bsum;
      } else {
#pragma unroll
        for (int c = 0; c < KW; c++) wacc[1][c] += acc[c];
        bacc[1] += bsum;
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) ptx::mbar_arrive(empty_bar(half));
  }
  // ---- the four row quarters of this CTA, added in fixed order; one partial [CO][KK + 1] per CTA (column KK = bias)
  if (a.part) {
    if (w_on) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        float* rp = red + ((wrq * 2 * CH + h * CH + wcl) * KH + wkh) * (KW + 1);
#pragma unroll
        for (int c = 0; c < KW; c++) rp[c] = wacc[h][c];
        rp[KW] = bacc[h];
      }
    }
    __syncthreads();
    if (w_on && wrq == 0) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int co = h * CH + wcl;
        if (co < a.CO) {
          float sum[KW + 1];
#pragma unroll
          for (int c = 0; c <= KW; c++) sum[c] = 0.f;
          for (int q = 0; q < RQ; q++) {
            const float* rp = red + ((q * 2 * CH + h * CH + wcl) * KH + wkh) * (KW + 1);
#pragma unroll
            for (int c = 0; c <= KW; c++) sum[c] = __fadd_rn(sum[c], rp[c]);
          }
          float* p = a.part + ((int64_t)blockIdx.x * a.CO + co) * (KK + 1);
#pragma unroll
          for (int c = 0; c < KW; c++) p[wkh * KW + c] = sum[c];
          if (wkh == 0) p[KK] = sum[KW];
        }
      }
    }
  }
}

// Fixed-order reduction of the per-CTA partials: block = 32 outputs x 8 slot groups; thread (o, zg) adds slots zg, zg+8, ...
// in ascending order (coalesced over o), the 8 group sums are added in order through shared memory.
__global__ void __launch_bounds__(256) c1_reduce_kernel(const float* __restrict__ part, int slots, int64_t Cout, int64_t Kc,
                                                        float* __restrict__ gk, float* __restrict__ gb) {
  __shared__ float sh[8][33];
  const int64_t Nv = Kc + 1, total = Cout * Nv;
  const int o = threadIdx.x & 31, zg = threadIdx.x >> 5;
  const int64_t idx = (int64_t)blockIdx.x * 32 + o;
  float acc = 0.f;
  if (idx < total)
    for (int z = zg; z < slots; z += 8) acc = __fadd_rn(acc, part[(int64_t)z * total + idx]);
  sh[zg][o] = acc;
  __syncthreads();
  if (zg == 0 && idx < total) {
    float s = sh[0][o];
#pragma unroll
    for (int q = 1; q < 8; q++) s = __fadd_rn(s, sh[q][o]);
    const int64_t co = idx / Nv, j = idx - co * Nv;
    if (j < Kc) { if (gk) gk[co * Kc + j] = s; }
    else if (gb) gb[co] = s;
  }
}

// ------------------------------------------------------------------ host side
static bool c1_shape_ok(const am_conv2d_desc& d, int64_t Wo) {
  // TMA staging: global row strides must be multiples of 16 bytes, and so must the byte offset of the box start along the
  // innermost dimension (-padW elements): padW % 4 == 0 (a misaligned start coordinate is an illegal instruction)
  return d.padW % 4 == 0 && d.C == 1 && d.strideH == 1 && d.strideW == 1 && d.dilH == 1 && d.dilW == 1 && d.kH == d.kW && (d.kH == 3 || d.kH == 5) &&
         d.Cout >= 1 && d.Cout <= 64 && d.H <= 64 && d.W <= 64 && d.padH < d.kH && d.padW < d.kW && d.W % 4 == 0 && Wo % 4 == 0 &&
         d.N < (1ll << 31);
}

int conv2d_forward_c1_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* input,
                          const float* kernel, const float* bias, float* output, int act, bool* done) {
  *done = false;
  if (!c1_shape_ok(d, Wo) || (reinterpret_cast<uintptr_t>(input) & 15) != 0 || !gemm_f32_tc_available()) return AM_OK;
  C1Args a{};
  a.w = kernel; a.bias = bias; a.y = output; a.N = d.N;
  a.H = (int)d.H; a.W = (int)d.W; a.CO = (int)d.Cout; a.padH = (int)d.padH; a.padW = (int)d.padW; a.HO = (int)Ho; a.WO = (int)Wo;
  a.GX = (a.WO + 3) / 4;
  a.XROWS = a.H + 2 * a.padH;
  const int KW = (int)d.kW;
  a.XP = ((4 * a.GX + KW - 1 + 3) / 4 * 4 > (a.W + 2 * a.padW + 3) / 4 * 4) ? (4 * a.GX + KW - 1 + 3) / 4 * 4 : (a.W + 2 * a.padW + 3) / 4 * 4;
  const int tasks = a.HO * a.GX;
  if (tasks > 320 || a.XP > 256 || a.XROWS > 256) return AM_OK;
  a.IMGS = 288 / tasks; if (a.IMGS < 1) a.IMGS = 1; if (a.IMGS > 8) a.IMGS = 8;
  const int threads = ((a.IMGS * tasks + 31) / 32) * 32;
  const int KKP = ((int)(d.kH * d.kW) + 3) / 4 * 4;
  const int buf_floats = (a.IMGS * a.XROWS * a.XP + 31) & ~31;
  const size_t smem = (size_t)(2 * buf_floats + a.CO * KKP + ((a.CO + 3) & ~3)) * 4 + 32 + 128;
  if (smem > 100 * 1024) return AM_OK;
  CUtensorMap tmX;
  const uint64_t dims[3] = {(uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.N};
  const uint64_t strides[2] = {(uint64_t)a.W * 4, (uint64_t)a.W * a.H * 4};
  const uint32_t box[3] = {(uint32_t)a.XP, (uint32_t)a.XROWS, (uint32_t)a.IMGS};
  int rc = make_tmap_f32_nd(&tmX, input, 3, dims, strides, box);
  if (rc) return rc;
  const int64_t ngroups = (a.N + a.IMGS - 1) / a.IMGS;
  auto kern = d.kH == 5 ? (act ? conv_c1_forward_kernel<5, 5, true> : conv_c1_forward_kernel<5, 5, false>)
                        : (act ? conv_c1_forward_kernel<3, 3, true> : conv_c1_forward_kernel<3, 3, false>);
  if (smem > 48 * 1024) AM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  AM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) return AM_OK;
  int64_t grid = (int64_t)per_sm * sm_count();              // persistent: every CTA resident, groups strided over them
  if (grid > ngroups) grid = ngroups;
  kern<<<(unsigned)grid, threads, smem, st>>>(tmX, a);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  *done = true;
  return AM_OK;
}

// Fused backward: grad_input, grad_kernel and grad_bias are final when *done (the partial reduction is launched here).
int conv2d_backward_c1_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* input,
                           const float* kernel, const float* grad_output, float* grad_input, float* grad_kernel,
                           float* grad_bias, bool* done) {
  *done = false;
  const bool want_gw = grad_kernel || grad_bias;
  if (!c1_shape_ok(d, Wo) || (want_gw && !input) || (grad_input && !kernel) || !gemm_f32_tc_available()) return AM_OK;
  if ((reinterpret_cast<uintptr_t>(grad_output) & 15) != 0 || (want_gw && (reinterpret_cast<uintptr_t>(input) & 15) != 0)) return AM_OK;
  C1Args a{};
  a.w = kernel; a.gi = grad_input; a.N = d.N;
  a.H = (int)d.H; a.W = (int)d.W; a.CO = (int)d.Cout; a.padH = (int)d.padH; a.padW = (int)d.padW; a.HO = (int)Ho; a.WO = (int)Wo;
  a.want_gi = grad_input != nullptr; a.want_gw = want_gw ? 1 : 0;
  const int KH = (int)d.kH, KW = (int)d.kW, KK = KH * KW, KWP = (KW + 3) / 4 * 4, LPAD = (KW - 1 + 3) / 4 * 4, RQ = 4;
  a.GX = (a.W + 3) / 4;
  a.XROWS = a.H + 2 * a.padH;
  // the weight-gradient threads read 4 floats past column WO + KW - 1 of an input row (next window prefetch)
  a.XP = ((a.W + 2 * a.padW + 3) / 4 * 4) + 8;
  if (a.XP < ((a.WO + 3) / 4 * 4) + 8) a.XP = ((a.WO + 3) / 4 * 4) + 8;
  // staged grad_output row: LPAD zero columns | WO values | zeros up to the last window any data-gradient thread reads
  // (furthest window: aligned base <= LPAD + 4*(GX-1) + padW - (KW-1), then up to 12 floats)
  a.GP = ((LPAD + 4 * (a.GX - 1) + a.padW - (KW - 1) + 12 + 3) / 4) * 4;
  if (a.GP < LPAD + ((a.WO + 3) / 4 * 4) + 4) a.GP = LPAD + ((a.WO + 3) / 4 * 4) + 4;
  // rows of a staged plane: HO plus a few zero rows so that consecutive planes start on different bank groups
  a.GROWS = a.HO;
  for (int extra = 0; extra < 8; extra++)
    if ((((a.HO + extra) * a.GP) & 31) % 8 == 4) { a.GROWS = a.HO + extra; break; }
  if (a.XP > 256 || a.XROWS > 256 || a.GP > 256 || a.GROWS > 256) return AM_OK;
  const int CH = (a.CO + 1) / 2;
  const int dthreads = a.want_gi ? ((a.H * a.GX + 31) / 32) * 32 : 0;
  const int wthreads = a.want_gw ? ((CH * KH * RQ + 31) / 32) * 32 : 0;
  int threads = dthreads + wthreads;
  if (threads < 128) threads = 128;
  if (threads > 448) return AM_OK;
  const int x_floats = (a.XROWS * a.XP + 31) & ~31, g_floats = (CH * a.GROWS * a.GP + 31) & ~31;
  const size_t smem = (size_t)(2 * x_floats + 2 * g_floats + KH * a.CO * KWP + RQ * 2 * CH * KH * (KW + 1)) * 4 + 32 + 128;
  if (smem > 112 * 1024) return AM_OK;
  CUtensorMap tmX, tmG;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)a.WO, (uint64_t)a.HO, (uint64_t)a.CO, (uint64_t)a.N};
    const uint64_t strides[3] = {(uint64_t)a.WO * 4, (uint64_t)a.WO * a.HO * 4, (uint64_t)a.WO * a.HO * a.CO * 4};
    const uint32_t box[4] = {(uint32_t)a.GP, (uint32_t)a.GROWS, (uint32_t)CH, 1u};
    if ((rc = make_tmap_f32_nd(&tmG, grad_output, 4, dims, strides, box))) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.N};
    const uint64_t strides[2] = {(uint64_t)a.W * 4, (uint64_t)a.W * a.H * 4};
    const uint32_t box[3] = {(uint32_t)a.XP, (uint32_t)a.XROWS, 1u};
    if ((rc = make_tmap_f32_nd(&tmX, want_gw ? input : grad_output, 3, dims, strides, box))) return rc;
  }
  auto kern = KH == 5 ? conv_c1_backward_kernel<5, 5> : conv_c1_backward_kernel<3, 3>;
  AM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  AM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) return AM_OK;
  int64_t grid = (int64_t)per_sm * sm_count();
  if (grid > a.N) grid = a.N;
  float* part = nullptr;
  if (want_gw) {
    void* p = nullptr;
    rc = workspace(kWsConv, (size_t)(grid * a.CO * (KK + 1)) * sizeof(float), &p);
    if (rc) return rc;
    part = (float*)p;                                   // every slot is written: each CTA has >= 1 image (grid <= N)
  }
  a.part = part;
  kern<<<(unsigned)grid, threads, smem, st>>>(tmX, tmG, a, dthreads, CH);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  if (want_gw) {
    const int64_t total = (int64_t)a.CO * (KK + 1);
    c1_reduce_kernel<<<(unsigned)ceil_div(total, 32), 256, 0, st>>>(part, (int)grid, a.CO, KK, grad_kernel, grad_bias);
    g_launch_count++;
    AM_CUDA_TRY(cudaGetLastError());
  }
  *done = true;
  return AM_OK;
}

}  // namespace am
