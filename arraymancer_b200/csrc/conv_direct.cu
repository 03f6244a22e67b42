// Direct float32 convolution with shared-memory-staged tiles and a register sliding window — the fast path of
// conv2d forward and of the stride-1 data gradient (LeNet-class shapes: kW in {1,3,5,7}, stride 1, dilation 1).
// Still an implicit GEMM (out[co, p] = sum_k W[co, k] * im2col[k, p]) and the im2col matrix still never exists:
//   * a CTA copies the raw NCHW input region of its images (zero halo = padding) and the matching weight slice
//     into shared memory, one block of input channels at a time;
//   * each thread owns CT output channels x PX consecutive output pixels of one row; per (ci, kh) it loads the
//     PX + KW - 1 input values it needs ONCE (128-bit LDS) and slides the KW taps over them in registers:
//     KW*CT*PX FMAs for ~(PX + KW - 1)/4 + KW*CT/4 shared loads, no per-element index decode or bounds test.
// ncu on the first (gather) version showed the FMA pipe 25-38 % busy with ~15 integer instructions per gathered
// element (profiles/r01_bringup.md §4); this kernel removes them.  Everything else — other dtypes, strided or
// dilated taps, tiles that do not fit — stays on the generic gather kernels of conv.cu.
//
// The data gradient of a stride-1 convolution is itself a convolution of grad_output with the transposed,
// spatially flipped weights and padding (k-1) - p (col2im of conv.nim:56-79 in gather form), so it runs through
// the same kernel with a strided view of the weight tensor.
#include <cstdlib>

#include "am_common.cuh"
#include "gemm_dispatch.h"

namespace am {

struct DirectArgs {
  const float* x;    // [N][C][H][W]
  const float* w;    // weight element (co, ci, kh, kw) at w[w_off + co*w_sco + ci*w_sci + kh*w_skh + kw*w_skw]
  const float* wp;   // the same weights packed by conv_pack_weights_kernel: wp[((ci*kH + kh)*KW + kw)*CO_P + co], zero padded
  int CO_P;          // padded channel count of wp (multiple of CO_B)
  const float* bias; // [CO] or null
  float* y;          // [N][CO][HO][WO]
  int64_t N;
  int C, H, W, CO, kH, padH, padW, HO, WO;
  int64_t w_off, w_sco, w_sci, w_skh, w_skw;
  // tiling
  int IMGS, TH, IH_T, IW_S;     // images / output rows per CTA; staged input rows, row pitch (multiple of 4)
  int SEGS;                     // PX-wide segments per output row = ceil(WO / PX)
  int CO_B, CG, CI_B;           // output channels per CTA (padded), channel groups, input channels per block
  int bands;                    // ceil(HO / TH)
  int BUF_FLOATS;               // floats per half of the double-buffered stage (multiple of 4)
  int relu;                     // fused activation: y = max(0, conv + bias) (nnp_activation.nim:35-36 semantics)
};

template <int CT, int PX, int KW>
__global__ void __launch_bounds__(512)
conv_direct_f32_kernel(const DirectArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* w_s = reinterpret_cast<float*>(smem_raw);                 // [CI_B][kH][KW][CO_B], co contiguous
  // two halves of [ weights [CI_B][kH][KW][CO_B] | inputs [IMGS][CI_B][IH_T][IW_S] ] (+ slack)
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t n0 = (int64_t)(blockIdx.x / a.bands) * a.IMGS;
  const int ho0 = (int)(blockIdx.x % a.bands) * a.TH;
  const int co0 = blockIdx.y * a.CO_B;
  const int imgs = (int)((a.N - n0 < a.IMGS) ? a.N - n0 : a.IMGS);
  const int th = (a.HO - ho0 < a.TH) ? a.HO - ho0 : a.TH;

  // thread -> (channel group, segment): segments enumerate (image, output row, PX-wide column block)
  const int nseg = imgs * th * a.SEGS;
  const int cg = tid % a.CG, sid = tid / a.CG;          // channel group fastest: a warp touches few windows
  const bool active = sid < nseg;
  int img = 0, row = 0, seg = 0;
  if (active) { img = sid / (th * a.SEGS); const int r = sid - img * th * a.SEGS; row = r / a.SEGS; seg = r - row * a.SEGS; }
  const int ch_stride = a.IH_T * a.IW_S;
  const int img_stride = a.CI_B * ch_stride;
  const int xbase = img * img_stride + row * a.IW_S + seg * PX;     // tap row kh adds kh * IW_S

  float acc[CT][PX];
#pragma unroll
  for (int c = 0; c < CT; c++)
#pragma unroll
    for (int j = 0; j < PX; j++) acc[c][j] = 0.f;

  const int h_base = ho0 - a.padH;
  const int wk_elems = a.kH * KW * a.CO_B;               // weights per input channel
  // staging geometry of a lane inside a plane: narrow rows are packed several per warp pass
  const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
  const int cols_per_iter = a.IW_S < 32 ? a.IW_S : 32;
  const int rows_per_iter = 32 / cols_per_iter;
  const int lane_row = lane / cols_per_iter, lane_col = lane - lane_row * cols_per_iter;
  const bool lane_ok = lane_row < rows_per_iter;
  // Double-buffered staging with cp.async (LDGSTS): while the FMAs consume channel block cb from one half of the
  // shared buffer, block cb+1 streams into the other half (zero-fill = padding) — one barrier per block instead of
  // two, and the staging latency hides under compute (ncu: CTA-barrier stalls were 18-53 % on the first version).
  const int buf_floats = a.BUF_FLOATS;                     // floats per buffer half (weights + inputs of one block)
  const uint32_t smem_u32base = (uint32_t)__cvta_generic_to_shared(w_s);
  auto issue_block = [&](int cb, int half) {
    const int cib = (a.C - cb < a.CI_B) ? a.C - cb : a.CI_B;
    const uint32_t wdst = smem_u32base + (uint32_t)(half * buf_floats) * 4u;
    const uint32_t xdst = wdst + (uint32_t)(a.CI_B * wk_elems) * 4u;
    // weight slice from the packed copy: rows of CO_B contiguous floats, 16-byte cp.async
    const int rows = cib * a.kH * KW, vec_per_row = a.CO_B / 4;
    const float* wsrc = a.wp + (size_t)cb * a.kH * KW * a.CO_P + co0;
    for (int i = tid; i < rows * vec_per_row; i += nt) {
      const int k = i / vec_per_row, v = i - k * vec_per_row;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wdst + 16u * (uint32_t)i), "l"(wsrc + (size_t)k * a.CO_P + 4 * v) : "memory");
    }
    // input region: one warp per (image, channel) plane, lanes over (rows_per_iter x IW_S) pixels, 4-byte cp.async
    // with zero fill outside the image — no per-element division
    for (int pl = warp; pl < imgs * cib; pl += nwarps) {
      const int im = pl / cib, c = pl - im * cib;
      const float* src = a.x + ((n0 + im) * a.C + cb + c) * (int64_t)a.H * a.W;
      const uint32_t dst = xdst + (uint32_t)(im * img_stride + c * ch_stride) * 4u;
      if (lane_ok) {
        for (int ih = lane_row; ih < a.IH_T; ih += rows_per_iter) {
          for (int iw = lane_col; iw < a.IW_S; iw += cols_per_iter) {
            const int h = h_base + ih, w = iw - a.padW;
            const bool ok = (unsigned)h < (unsigned)a.H && (unsigned)w < (unsigned)a.W;
            const float* sp = ok ? src + h * a.W + w : a.x;
            const int nbytes = ok ? 4 : 0;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + (uint32_t)(ih * a.IW_S + iw) * 4u), "l"(sp), "r"(nbytes) : "memory");
          }
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue_block(0, 0);
  int half = 0;
  for (int cb = 0; cb < a.C; cb += a.CI_B, half ^= 1) {
    const int cib = (a.C - cb < a.CI_B) ? a.C - cb : a.CI_B;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                      // block cb landed; everyone is done with the other half
    if (cb + a.CI_B < a.C) issue_block(cb + a.CI_B, half ^ 1);
    const float* w_cur = w_s + half * buf_floats;
    const float* in_cur = w_cur + a.CI_B * wk_elems;
    if (active) {
      for (int ci = 0; ci < cib; ci++) {
        const float* xr = in_cur + xbase + ci * ch_stride;
        const float* wr = w_cur + ci * wk_elems + cg * CT;
        for (int kh = 0; kh < a.kH; kh++) {
          // PX + KW - 1 consecutive inputs of tap row kh, as 128-bit loads (xbase and IW_S are multiples of 4)
          constexpr int NX = (PX + KW - 1 + 3) / 4;
          union { int4 q[NX]; float e[NX * 4]; } xw;
#pragma unroll
          for (int v = 0; v < NX; v++) xw.q[v] = reinterpret_cast<const int4*>(xr)[v];
          xr += a.IW_S;
#pragma unroll
          for (int kw = 0; kw < KW; kw++) {
            union { int4 q[CT / 4]; float e[CT]; } wv;
#pragma unroll
            for (int g = 0; g < CT / 4; g++) wv.q[g] = reinterpret_cast<const int4*>(wr)[g];
            wr += a.CO_B;
#pragma unroll
            for (int j = 0; j < PX; j++)
#pragma unroll
              for (int c = 0; c < CT; c++) acc[c][j] = fmaf(wv.e[c], xw.e[j + kw], acc[c][j]);
          }
        }
      }
    }
  }

  // ---- epilogue: + bias (and the fused ReLU), NCHW stores
  const bool relu = a.relu != 0;
  auto act = [relu](float v) { return (relu && v <= 0.f) ? 0.f : v; };
  if (active) {
    const int wo0 = seg * PX;
    const int64_t chan_pix = (int64_t)a.HO * a.WO;
    float* ypix = a.y + (((n0 + img) * a.CO) * (int64_t)a.HO + (ho0 + row)) * a.WO + wo0;
#pragma unroll
    for (int c = 0; c < CT; c++) {
      const int co = co0 + cg * CT + c;
      if (co < a.CO) {
        const float b = a.bias ? a.bias[co] : 0.f;
        float* yc = ypix + co * chan_pix;
        if (wo0 + PX <= a.WO && ((reinterpret_cast<uintptr_t>(yc) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < PX; j += 4)
            *reinterpret_cast<float4*>(yc + j) = make_float4(act(__fadd_rn(acc[c][j], b)), act(__fadd_rn(acc[c][j + 1], b)),
                                                             act(__fadd_rn(acc[c][j + 2], b)), act(__fadd_rn(acc[c][j + 3], b)));
        } else {
#pragma unroll
          for (int j = 0; j < PX; j++)
            if (wo0 + j < a.WO) yc[j] = act(__fadd_rn(acc[c][j], b));
        }
      }
    }
  }
}

// wp[k][c] = W(c, ci, kh, kw) for k = (ci*kH + kh)*KW + kw, c < CO_P (zero for c >= CO): one tiny pass per call,
// so every CTA's weight staging is a contiguous 128-bit copy instead of a strided gather.
__global__ void conv_pack_weights_kernel(const float* __restrict__ w, float* __restrict__ wp, int CO, int CO_P, int C,
                                         int kH, int KW, int64_t w_off, int64_t w_sco, int64_t w_sci, int64_t w_skh,
                                         int64_t w_skw) {
  const int total = C * kH * KW * CO_P;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % CO_P, k = i / CO_P;
    const int ci = k / (kH * KW), r = k - ci * (kH * KW), kh = r / KW, kw = r - kh * KW;
    wp[i] = (c < CO) ? w[w_off + c * w_sco + ci * w_sci + kh * w_skh + kw * w_skw] : 0.f;
  }
}

// Picks a tiling; returns false when the shape does not fit the fast path (caller falls back to the gather kernel).
static bool plan_direct(DirectArgs& a, int KW, int* ct_out, int* px_out, int* nthreads, size_t* smem, int* grid_y) {
  auto padded = [](int v, int m) { return (v + m - 1) / m * m; };
  // register tile: 4 channels x 8 (or 4) pixels (65 / 50 registers).  Chosen by a sweep over CT in {4,8}, PX in
  // {4,8}, 128..512 threads and 24..80 KB of staging on the LeNet layers (tools/gpu_bringup.py conv_sweep,
  // profiles/r01_bringup.md §4): the 4-channel tile won everywhere (more CTAs per SM hide the staging phases);
  // deep reductions (K' >= 256) want 512-thread CTAs (weight staging amortised), shallow ones 128.
  auto envi = [](const char*, int dflt) { return dflt; };   // (the round-1 sweep overrides are gone: the library reads no environment)
  const int CT = envi("AM_CONV_CT", 4);
  const int PX = envi("AM_CONV_PX", (a.WO % 8 == 0 || a.WO >= 32) ? 8 : 4);
  int CO_B = padded(a.CO, CT);
  if (CO_B > 64) CO_B = 64;                               // more output channels -> grid.y
  const int CG = CO_B / CT;
  const int SEGS = (a.WO + PX - 1) / PX;
  const int max_threads = envi("AM_CONV_MAXT", (a.C * a.kH * KW >= 256) ? 512 : 128);
  const int max_segs = max_threads / CG;                  // segments one CTA can hold
  if (max_segs < 1 || SEGS > max_segs) return false;      // one output row must fit
  int TH = a.HO, IMGS = 1;
  if (a.HO * SEGS <= max_segs) {
    IMGS = max_segs / (a.HO * SEGS);
    const int64_t by_grid = a.N / (8 * (int64_t)sm_count());     // keep the grid several waves deep
    if (IMGS > by_grid) IMGS = (int)(by_grid > 0 ? by_grid : 1);
    if (IMGS > 16) IMGS = 16;
    if (envi("AM_CONV_IMGS", 0) > 0 && envi("AM_CONV_IMGS", 0) <= max_segs / (a.HO * SEGS)) IMGS = envi("AM_CONV_IMGS", 0);
  } else {
    const int th_max = max_segs / SEGS;
    const int nb = (a.HO + th_max - 1) / th_max;
    TH = (a.HO + nb - 1) / nb;                            // balanced row bands
  }
  const int IW_T = (a.WO - 1) + (KW - 1) + 1;
  // row pitch: multiple of 4 (128-bit loads) and wide enough for the last segment's full window
  const int IW_S = padded(SEGS * PX + KW - 1 > IW_T ? SEGS * PX + KW - 1 : IW_T, 4);
  const int IH_T = (TH - 1) + (a.kH - 1) + 1;
  // input-channel block: keep weights + inputs of a block under ~40 KB so several CTAs share an SM
  const size_t per_ci = ((size_t)a.kH * KW * CO_B + (size_t)IMGS * IH_T * IW_S) * sizeof(float);
  int CI_B = (int)(((size_t)envi("AM_CONV_SMEMKB", 40) * 1024) / per_ci);    // per half of the double buffer
  if (CI_B < 1) {
    if (per_ci > 100 * 1024) return false;          // two halves must fit the 227 KB of an SM
    CI_B = 1;
  }
  if (CI_B > a.C) CI_B = a.C;
  a.IMGS = IMGS; a.TH = TH; a.IH_T = IH_T; a.IW_S = IW_S; a.SEGS = SEGS; a.CO_B = CO_B; a.CG = CG; a.CI_B = CI_B;
  a.bands = (a.HO + TH - 1) / TH;
  *ct_out = CT; *px_out = PX;
  const int segs_cta = IMGS * TH * SEGS;
  *nthreads = padded(segs_cta * CG, 32);
  if (*nthreads > 512) return false;
  a.BUF_FLOATS = (int)padded((int)(((size_t)CI_B * per_ci) / sizeof(float)) + 16, 4);   // + slack: the last window may read a few floats past the tile
  *smem = 2 * (size_t)a.BUF_FLOATS * sizeof(float);
  *grid_y = (a.CO + CO_B - 1) / CO_B;
  return true;
}

template <int CT, int PX>
static cudaError_t launch_kw(int KW, dim3 grid, int nthreads, size_t smem, cudaStream_t st, const DirectArgs& a) {
  auto go = [&](auto kern) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, nthreads, smem, st>>>(a);
    return cudaGetLastError();
  };
  switch (KW) {
    case 1: return go(conv_direct_f32_kernel<CT, PX, 1>);
    case 3: return go(conv_direct_f32_kernel<CT, PX, 3>);
    case 5: return go(conv_direct_f32_kernel<CT, PX, 5>);
    default: return go(conv_direct_f32_kernel<CT, PX, 7>);
  }
}

static int launch_direct(cudaStream_t st, DirectArgs& a, int KW, bool* done) {
  *done = false;
  if (KW != 1 && KW != 3 && KW != 5 && KW != 7) return AM_OK;
  int ct = 0, px = 0, nthreads = 0, grid_y = 0;
  size_t smem = 0;
  if (!plan_direct(a, KW, &ct, &px, &nthreads, &smem, &grid_y)) return AM_OK;
  const int64_t gx = ceil_div(a.N, a.IMGS) * a.bands;
  if (gx > 2147483647ll || grid_y > 65535) return AM_OK;
  dim3 grid((unsigned)gx, (unsigned)grid_y);
  // pack the weights once (k-major, channel contiguous, padded to grid_y * CO_B channels)
  a.CO_P = grid_y * a.CO_B;
  const int64_t wtotal = (int64_t)a.C * a.kH * KW * a.CO_P;
  if (wtotal >= (1ll << 31)) return AM_OK;
  void* wp = nullptr;
  int rcw = workspace(kWsConvW, (size_t)wtotal * sizeof(float), &wp);
  if (rcw) return rcw;
  a.wp = (const float*)wp;
  conv_pack_weights_kernel<<<(unsigned)ceil_div(wtotal, 256), 256, 0, st>>>(a.w, (float*)wp, a.CO, a.CO_P, a.C, a.kH, KW,
                                                                            a.w_off, a.w_sco, a.w_sci, a.w_skh, a.w_skw);
  g_launch_count++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "conv_pack_weights launch");
  if (ct == 8 && px == 8) e = launch_kw<8, 8>(KW, grid, nthreads, smem, st, a);
  else if (ct == 8) e = launch_kw<8, 4>(KW, grid, nthreads, smem, st, a);
  else if (px == 8) e = launch_kw<4, 8>(KW, grid, nthreads, smem, st, a);
  else e = launch_kw<4, 4>(KW, grid, nthreads, smem, st, a);
  if (e != cudaSuccess) return cuda_fail(e, "conv_direct launch");
  g_launch_count++;
  *done = true;
  return AM_OK;
}

// forward: y = conv(x, w) + bias.  *done = false -> caller must use the gather kernels.
int conv2d_forward_direct_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* input,
                              const float* kernel, const float* bias, float* output, int act, bool* done) {
  *done = false;
  if (d.strideH != 1 || d.strideW != 1 || d.dilH != 1 || d.dilW != 1) return AM_OK;
  DirectArgs a{};
  a.relu = act;
  a.x = input; a.w = kernel; a.bias = bias; a.y = output; a.N = d.N;
  a.C = (int)d.C; a.H = (int)d.H; a.W = (int)d.W; a.CO = (int)d.Cout; a.kH = (int)d.kH;
  a.padH = (int)d.padH; a.padW = (int)d.padW; a.HO = (int)Ho; a.WO = (int)Wo;
  a.w_off = 0; a.w_sco = d.C * d.kH * d.kW; a.w_sci = d.kH * d.kW; a.w_skh = d.kW; a.w_skw = 1;
  return launch_direct(st, a, (int)d.kW, done);
}

// data gradient for stride 1, dilation 1: grad_input = conv(grad_output, flip(W)^T) with padding (k-1) - p.
int conv2d_dgrad_direct_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* grad_output,
                            const float* kernel, float* grad_input, bool* done) {
  *done = false;
  if (d.strideH != 1 || d.strideW != 1 || d.dilH != 1 || d.dilW != 1) return AM_OK;
  DirectArgs a{};
  a.x = grad_output; a.w = kernel; a.bias = nullptr; a.y = grad_input; a.N = d.N;
  a.C = (int)d.Cout; a.H = (int)Ho; a.W = (int)Wo; a.CO = (int)d.C; a.kH = (int)d.kH;
  a.padH = (int)((d.kH - 1) - d.padH); a.padW = (int)((d.kW - 1) - d.padW);
  a.HO = (int)d.H; a.WO = (int)d.W;
  // "output channel" = ci of the original weights, "input channel" = co, taps flipped
  a.w_sco = d.kH * d.kW; a.w_sci = d.C * d.kH * d.kW; a.w_skh = -d.kW; a.w_skw = -1;
  a.w_off = (d.kH - 1) * d.kW + (d.kW - 1);
  return launch_direct(st, a, (int)d.kW, done);
}


// =====================================================================================================
// Weight gradient (+ bias gradient), float32, stride 1, dilation 1, kW in {1,3,5,7}:
//   gW[co][ci][kh][kw] = sum_{n,ho,wo} gout[n,co,ho,wo] * x[n,ci,ho-pH+kh,wo-pW+kw]      (conv.nim:140)
//   gb[co]             = sum_{n,ho,wo} gout[n,co,ho,wo]                                   (nnp_convolution.nim:94)
// A CTA walks a contiguous range of images in batches staged in shared memory.  A thread owns CT output
// channels x one (ci, kh) tap row x all KW taps (CT*KW accumulators) and a slice of the output-pixel segments:
// per PX-wide segment it loads CT gout windows and ONE input window and slides the taps in registers
// (CT*KW*PX FMAs per ~CT*PX/4 + (PX+KW-1)/4 shared loads).  Slices are summed in shared memory in a fixed order,
// CTAs write partials, and the existing fixed-order second pass (wgrad_reduce_kernel) adds them: deterministic.
struct WgradArgs {
  const float* x;     // [N][C][H][W]
  const float* g;     // [N][CO][HO][WO]
  float* part;        // [groups][CO][Kc + 1]
  int64_t N;
  int C, H, W, CO, kH, padH, padW, HO, WO;
  int IMGS, LOOPS;    // images per staged batch, batches per CTA
  int IH, IW_S, GW_S; // staged input rows / pitch, staged gout row pitch (multiples of 4)
  int X_CH;           // staged input plane pitch: >= IH*IW_S, multiple of 4 and == 4 (mod 32)
  int SEGS;           // PX-wide segments per output row
  int CO_B, CG, CI_B; // channels per CTA
  int UNITS, SLICES;  // (cg, ci, kh) units per CTA, pixel slices per unit
  int Kc;             // C*kH*KW
};

template <int CT, int PX, int KW>
__global__ void __launch_bounds__(512)
conv_wgrad_direct_f32_kernel(const WgradArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* g_s = reinterpret_cast<float*>(smem_raw);                         // [IMGS][CO_B][HO][GW_S]
  float* x_s = g_s + (size_t)a.IMGS * a.CO_B * a.HO * a.GW_S;              // [IMGS][CI_B][IH][IW_S]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int co0 = blockIdx.y * a.CO_B, ci0 = blockIdx.z * a.CI_B;
  const int cib = (a.C - ci0 < a.CI_B) ? a.C - ci0 : a.CI_B;
  const int64_t nbeg = (int64_t)blockIdx.x * a.IMGS * a.LOOPS;

  // thread -> (unit = (cg, ci, kh), slice)
  // lanes run over ci (x windows land in different banks: the plane pitch is 4 mod 32), then kh, then cg
  // (gout windows are then warp-broadcasts; with cg fastest they were 8-way bank conflicts: planes are 64 floats apart)
  const int unit = tid % a.UNITS, slice = tid / a.UNITS;
  const int ci = unit % a.CI_B, r = unit / a.CI_B, kh = r % a.kH, cg = r / a.kH;
  const bool active = slice < a.SLICES && ci < cib;
  const bool bias_unit = (ci == 0 && kh == 0 && ci0 == 0);

  float acc[CT][KW];
  float accb[CT];
#pragma unroll
  for (int c = 0; c < CT; c++) {
    accb[c] = 0.f;
#pragma unroll
    for (int k = 0; k < KW; k++) acc[c][k] = 0.f;
  }
  const int g_ch = a.HO * a.GW_S, g_img = a.CO_B * g_ch;
  const int x_ch = a.X_CH, x_img = a.CI_B * x_ch;
  // staging geometry of a lane inside a plane (narrow rows are packed several per warp pass)
  const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
  const int gcols_per_iter = a.GW_S < 32 ? a.GW_S : 32, grows_per_iter = 32 / gcols_per_iter;
  const int glane_row = lane / gcols_per_iter, glane_col = lane - glane_row * gcols_per_iter;
  const bool glane_ok = glane_row < grows_per_iter;
  const int xcols_per_iter = a.IW_S < 32 ? a.IW_S : 32, xrows_per_iter = 32 / xcols_per_iter;
  const int xlane_row = lane / xcols_per_iter, xlane_col = lane - xlane_row * xcols_per_iter;
  const bool xlane_ok = xlane_row < xrows_per_iter;

  for (int lp = 0; lp < a.LOOPS; lp++) {
    const int64_t n0 = nbeg + (int64_t)lp * a.IMGS;
    if (n0 >= a.N) break;
    const int imgs = (int)((a.N - n0 < a.IMGS) ? a.N - n0 : a.IMGS);
    __syncthreads();
    // ---- stage gout: one warp per (image, channel) plane; a plane is contiguous in global memory
    for (int pl = warp; pl < imgs * a.CO_B; pl += nwarps) {
      const int im = pl / a.CO_B, c = pl - im * a.CO_B;
      float* dst = g_s + im * g_img + c * g_ch;
      if (co0 + c < a.CO) {
        const float* src = a.g + ((n0 + im) * a.CO + co0 + c) * (int64_t)a.HO * a.WO;
        if (a.GW_S == a.WO) {
          for (int e = lane; e < g_ch; e += 32) dst[e] = src[e];
        } else if (glane_ok) {
          for (int ho = glane_row; ho < a.HO; ho += grows_per_iter)
            for (int wo = glane_col; wo < a.GW_S; wo += gcols_per_iter) dst[ho * a.GW_S + wo] = (wo < a.WO) ? src[ho * a.WO + wo] : 0.f;
        }
      } else {
        for (int e = lane; e < g_ch; e += 32) dst[e] = 0.f;
      }
    }
    // ---- stage the input block with its zero halo (columns past WO + KW - 2 are never multiplied by a non-zero
    //      gout value, but are zeroed too so that Inf/NaN in unused pixels cannot leak in as 0 * Inf)
    for (int pl = warp; pl < imgs * a.CI_B; pl += nwarps) {
      const int im = pl / a.CI_B, c = pl - im * a.CI_B;
      float* dst = x_s + im * x_img + c * x_ch;
      const float* src = a.x + ((n0 + im) * a.C + ci0 + c) * (int64_t)a.H * a.W;
      if (xlane_ok) {
        for (int ih = xlane_row; ih < a.IH; ih += xrows_per_iter)
          for (int iw = xlane_col; iw < a.IW_S; iw += xcols_per_iter) {
            const int h = ih - a.padH, w = iw - a.padW;
            float v = 0.f;
            if (c < cib && iw < a.WO + KW - 1 && (unsigned)h < (unsigned)a.H && (unsigned)w < (unsigned)a.W) v = src[h * a.W + w];
            dst[ih * a.IW_S + iw] = v;
          }
      }
    }
    __syncthreads();
    if (active) {
      // this thread's slice of the (image, row) pairs; all SEGS segments of a row stay together
      const int rows_total = imgs * a.HO;
      const int per = (rows_total + a.SLICES - 1) / a.SLICES;
      const int rbeg = slice * per, rend = (rbeg + per < rows_total) ? rbeg + per : rows_total;
      for (int rr = rbeg; rr < rend; rr++) {
        const int im = rr / a.HO, ho = rr - im * a.HO;
        const float* gp = g_s + im * g_img + (cg * CT) * g_ch + ho * a.GW_S;
        const float* xp = x_s + im * x_img + ci * x_ch + (ho + kh) * a.IW_S;
        for (int sg = 0; sg < a.SEGS; sg++) {
          constexpr int NX = (PX + KW - 1 + 3) / 4;
          union { int4 q[NX]; float e[NX * 4]; } xw;
#pragma unroll
          for (int v = 0; v < NX; v++) xw.q[v] = reinterpret_cast<const int4*>(xp + sg * PX)[v];
#pragma unroll
          for (int c = 0; c < CT; c++) {
            union { int4 q[PX / 4]; float e[PX]; } gw;
#pragma unroll
            for (int v = 0; v < PX / 4; v++) gw.q[v] = reinterpret_cast<const int4*>(gp + c * g_ch + sg * PX)[v];
#pragma unroll
            for (int j = 0; j < PX; j++) {
#pragma unroll
              for (int k = 0; k < KW; k++) acc[c][k] = fmaf(gw.e[j], xw.e[j + k], acc[c][k]);
            }
            if (bias_unit) {
#pragma unroll
              for (int j = 0; j < PX; j++) accb[c] += gw.e[j];
            }
          }
        }
      }
    }
  }

  // ---- fixed-order reduction over the slices in shared memory, then one partial per CTA
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem_raw);          // [SLICES][UNITS][CT*(KW+1)]
  constexpr int PER = CT * (KW + 1);
  if (slice < a.SLICES) {
    float* dst = red + ((size_t)slice * a.UNITS + unit) * PER;
#pragma unroll
    for (int c = 0; c < CT; c++) {
#pragma unroll
      for (int k = 0; k < KW; k++) dst[c * (KW + 1) + k] = acc[c][k];
      dst[c * (KW + 1) + KW] = accb[c];
    }
  }
  __syncthreads();
  const int Nv = a.Kc + 1;
  float* pout = a.part + (size_t)blockIdx.x * a.CO * Nv;
  for (int i = tid; i < a.UNITS * PER; i += nt) {
    const int u = i / PER, e = i - u * PER, c = e / (KW + 1), k = e - c * (KW + 1);
    const int uci = u % a.CI_B, ur = u / a.CI_B, ukh = ur % a.kH, ucg = ur / a.kH;
    const int co = co0 + ucg * CT + c;
    if (co >= a.CO || uci >= cib) continue;
    float s = 0.f;
    for (int sl = 0; sl < a.SLICES; sl++) s += red[((size_t)sl * a.UNITS + u) * PER + e];
    if (k < KW) pout[(size_t)co * Nv + ((ci0 + uci) * a.kH + ukh) * KW + k] = s;
    else if (uci == 0 && ukh == 0 && ci0 == 0) pout[(size_t)co * Nv + a.Kc] = s;
  }
}

static bool plan_wgrad(WgradArgs& a, int KW, int* ct_out, int* px_out, int* nthreads, size_t* smem, dim3* grid) {
  auto padded = [](int v, int m) { return (v + m - 1) / m * m; };
  const int CT = 4;
  const int PX = (a.WO % 8 == 0 || a.WO >= 32) ? 8 : 4;
  int CO_B = padded(a.CO, CT);
  if (CO_B > 32) CO_B = 32;
  const int CG = CO_B / CT;
  // units = CG * CI_B * kH <= 512
  int CI_B = 512 / (CG * a.kH);
  if (CI_B < 1) return false;
  if (CI_B > a.C) CI_B = a.C;
  const int UNITS = CG * CI_B * a.kH;
  int SLICES = 512 / UNITS;
  if (SLICES < 1) return false;
  const int SEGS = (a.WO + PX - 1) / PX;
  const int GW_S = padded(SEGS * PX, 4);
  const int IH = a.HO + a.kH - 1;
  const int IW_S = padded(SEGS * PX + KW - 1, 4);
  int X_CH = padded(IH * IW_S, 4);
  while (X_CH % 32 != 4) X_CH += 4;
  const size_t per_img = ((size_t)CO_B * a.HO * GW_S + (size_t)CI_B * X_CH) * sizeof(float);
  int IMGS = (int)((96 * 1024) / per_img);
  if (IMGS < 1) {
    if (per_img > 160 * 1024) return false;
    IMGS = 1;
  }
  if (IMGS > 32) IMGS = 32;
  if (SLICES > IMGS * a.HO) SLICES = IMGS * a.HO;
  const size_t red_bytes = (size_t)SLICES * UNITS * CT * (KW + 1) * sizeof(float);
  const size_t stage_bytes = (size_t)IMGS * per_img + 64;
  if (red_bytes > 200 * 1024) return false;
  // image groups: about 2 CTAs per SM in total over (groups x co blocks x ci blocks)
  const int yb = (a.CO + CO_B - 1) / CO_B, zb = (a.C + CI_B - 1) / CI_B;
  int64_t groups = (2 * (int64_t)sm_count() + yb * zb - 1) / (yb * zb);
  const int64_t batches = (a.N + IMGS - 1) / IMGS;
  if (groups > batches) groups = batches;
  if (groups < 1) groups = 1;
  const int LOOPS = (int)((batches + groups - 1) / groups);
  groups = (batches + LOOPS - 1) / LOOPS;
  if (groups > 65535 || yb > 65535 || zb > 65535) return false;
  a.IMGS = IMGS; a.LOOPS = LOOPS; a.IH = IH; a.IW_S = IW_S; a.GW_S = GW_S; a.SEGS = SEGS; a.X_CH = X_CH;
  a.CO_B = CO_B; a.CG = CG; a.CI_B = CI_B; a.UNITS = UNITS; a.SLICES = SLICES;
  *ct_out = CT; *px_out = PX;
  *nthreads = padded(UNITS * SLICES, 32);
  if (*nthreads > 512) return false;
  *smem = stage_bytes > red_bytes ? stage_bytes : red_bytes;
  *grid = dim3((unsigned)groups, (unsigned)yb, (unsigned)zb);
  return true;
}

// Weight + bias gradient partials; *groups_out = number of partial slabs written to `part`
// ([groups][CO][Kc+1]); the caller reduces them with wgrad_reduce_kernel.  *done = false -> fallback.
int conv2d_wgrad_direct_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* input,
                            const float* grad_output, float** part_out, int* groups_out, bool* done) {
  *done = false;
  if (d.strideH != 1 || d.strideW != 1 || d.dilH != 1 || d.dilW != 1) return AM_OK;
  const int KW = (int)d.kW;
  if (KW != 1 && KW != 3 && KW != 5 && KW != 7) return AM_OK;
  WgradArgs a{};
  a.x = input; a.g = grad_output; a.N = d.N;
  a.C = (int)d.C; a.H = (int)d.H; a.W = (int)d.W; a.CO = (int)d.Cout; a.kH = (int)d.kH;
  a.padH = (int)d.padH; a.padW = (int)d.padW; a.HO = (int)Ho; a.WO = (int)Wo;
  a.Kc = (int)(d.C * d.kH * d.kW);
  int ct = 0, px = 0, nthreads = 0;
  size_t smem = 0;
  dim3 grid;
  if (!plan_wgrad(a, KW, &ct, &px, &nthreads, &smem, &grid)) return AM_OK;
  void* part = nullptr;
  const size_t slab = (size_t)a.CO * (a.Kc + 1);
  int rc = workspace(kWsConv, (size_t)grid.x * slab * sizeof(float), &part);
  if (rc) return rc;
  a.part = (float*)part;
  // ci blocks / co blocks write disjoint entries of a slab, but not all of them when a CTA has no work: clear first
  AM_CUDA_TRY(cudaMemsetAsync(part, 0, (size_t)grid.x * slab * sizeof(float), st));
  auto go = [&](auto kern) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, nthreads, smem, st>>>(a);
    return cudaGetLastError();
  };
  cudaError_t e;
  if (px == 8) {
    switch (KW) {
      case 1: e = go(conv_wgrad_direct_f32_kernel<4, 8, 1>); break;
      case 3: e = go(conv_wgrad_direct_f32_kernel<4, 8, 3>); break;
      case 5: e = go(conv_wgrad_direct_f32_kernel<4, 8, 5>); break;
      default: e = go(conv_wgrad_direct_f32_kernel<4, 8, 7>); break;
    }
  } else {
    switch (KW) {
      case 1: e = go(conv_wgrad_direct_f32_kernel<4, 4, 1>); break;
      case 3: e = go(conv_wgrad_direct_f32_kernel<4, 4, 3>); break;
      case 5: e = go(conv_wgrad_direct_f32_kernel<4, 4, 5>); break;
      default: e = go(conv_wgrad_direct_f32_kernel<4, 4, 7>); break;
    }
  }
  if (e != cudaSuccess) return cuda_fail(e, "conv_wgrad_direct launch");
  g_launch_count++;
  *part_out = (float*)part;
  *groups_out = (int)grid.x;
  *done = true;
  return AM_OK;
}

}  // namespace am
