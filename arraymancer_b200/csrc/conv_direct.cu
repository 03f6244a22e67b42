// Direct convolution with the input tile staged in shared memory — the fast path of conv2d forward
// and of the stride-1 data gradient.  Still an implicit GEMM (out[co, p] = sum_k W[co, k] * im2col[k, p]) and the
// im2col matrix still never exists: a CTA copies the raw NCHW input region of a group of images (zero halo =
// padding) and its slice of the weights into shared memory ONCE, then every thread keeps a CO_T x PX_T register
// tile and walks k = (ci, kh, kw) with plain address increments — no per-element index decode, bounds test or
// table lookup as in the generic gather kernel (conv.cu), which stays as the fallback for shapes whose tiles do
// not fit in shared memory, and for strided dgrad / wgrad.
//
// The data gradient of a stride-1 convolution is itself a convolution of grad_output with the transposed,
// spatially flipped weights and padding d*(k-1) - p (col2im of conv.nim:56-79 in gather form), so it runs
// through the same kernel with a strided view of the weight tensor.
#include "am_common.cuh"
#include "gemm_dispatch.h"

namespace am {

template <class T>
struct DirectArgs {
  const T* x;        // [N][C][H][W]
  const T* w;        // weight element (co, ci, kh, kw) at w[w_off + co*w_sco + ci*w_sci + kh*w_skh + kw*w_skw]
  const T* bias;     // [CO] or null
  T* y;              // [N][CO][HO][WO]
  int64_t N;
  int C, H, W, CO, kH, kW, padH, padW, sH, sW, dH, dW, HO, WO;
  int64_t w_off, w_sco, w_sci, w_skh, w_skw;
  // tiling
  int IMGS, TH, IH_T, IW_T, IW_S;   // images / output rows per CTA; staged input rows, cols, padded row pitch
  int CO_B, CG, PG, KP;             // channels per CTA (padded), channel groups, pixel groups, K' = C*kH*kW
  int bands;                        // ceil(HO / TH)
};

template <class T, int CO_T, int PX_T>
__global__ void __launch_bounds__(512)
conv_direct_kernel(const DirectArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* w_s = reinterpret_cast<T*>(smem_raw);                       // [KP][CO_B], co contiguous
  T* in_s = w_s + (size_t)a.KP * a.CO_B;                         // [IMGS][C][IH_T][IW_S]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t n0 = (int64_t)(blockIdx.x / a.bands) * a.IMGS;
  const int ho0 = (int)(blockIdx.x % a.bands) * a.TH;
  const int co0 = blockIdx.y * a.CO_B;
  const int imgs = (int)((a.N - n0 < a.IMGS) ? a.N - n0 : a.IMGS);
  const int th = (a.HO - ho0 < a.TH) ? a.HO - ho0 : a.TH;

  // ---- stage the weight slice: w_s[k][c] = W(co0 + c, ci, kh, kw), zero for padded channels
  const int kHW = a.kH * a.kW;
  for (int i = tid; i < a.KP * a.CO_B; i += nt) {
    const int k = i / a.CO_B, c = i - k * a.CO_B;       // consecutive threads -> consecutive k would be the coalesced order
    const int ci = k / kHW, r = k - ci * kHW, kh = r / a.kW, kw = r - kh * a.kW;
    const int co = co0 + c;
    w_s[i] = (co < a.CO) ? a.w[a.w_off + co * a.w_sco + ci * a.w_sci + kh * a.w_skh + kw * a.w_skw] : T(0);
  }
  // ---- stage the input region (zero outside the image = padding)
  const int img_elems = a.C * a.IH_T * a.IW_S;
  const int h_base = ho0 * a.sH - a.padH;
  for (int i = tid; i < imgs * img_elems; i += nt) {
    const int img = i / img_elems, r0 = i - img * img_elems;
    const int c = r0 / (a.IH_T * a.IW_S), r1 = r0 - c * (a.IH_T * a.IW_S);
    const int ih = r1 / a.IW_S, iw = r1 - ih * a.IW_S;
    const int h = h_base + ih, w = iw - a.padW;
    T v = T(0);
    if (iw < a.IW_T && (unsigned)h < (unsigned)a.H && (unsigned)w < (unsigned)a.W)
      v = a.x[(((n0 + img) * a.C + c) * a.H + h) * (int64_t)a.W + w];
    in_s[i] = v;
  }
  __syncthreads();

  const int cg = tid / a.PG, pg = tid - cg * a.PG;
  const int ptile = imgs * th * a.WO;                 // valid output pixels of this CTA
  const int slots = a.PG * PX_T;
  const int ch_stride = a.IH_T * a.IW_S;
  const T* wbase = w_s + cg * CO_T;

  for (int q0 = 0; q0 < ptile; q0 += slots) {
    int boff[PX_T];
    int ooff[PX_T];                                   // relative to this CTA's first (image, channel 0, row ho0)
#pragma unroll
    for (int j = 0; j < PX_T; j++) {
      const int q = q0 + pg + j * a.PG;
      if (q < ptile) {
        const int img = q / (th * a.WO), r = q - img * (th * a.WO);
        const int row = r / a.WO, wo = r - row * a.WO;
        boff[j] = img * img_elems + row * a.sH * a.IW_S + wo * a.sW;
        ooff[j] = ((img * a.CO) * a.HO + row) * a.WO + wo;
      } else {
        boff[j] = 0;
        ooff[j] = -1;
      }
    }
    T acc[CO_T][PX_T];
#pragma unroll
    for (int c = 0; c < CO_T; c++)
#pragma unroll
      for (int j = 0; j < PX_T; j++) acc[c][j] = T(0);

    const T* wk = wbase;
    for (int ci = 0; ci < a.C; ci++) {
      for (int kh = 0; kh < a.kH; kh++) {
        const int rowoff = ci * ch_stride + kh * a.dH * a.IW_S;
        for (int kw = 0; kw < a.kW; kw++) {
          const int koff = rowoff + kw * a.dW;
          constexpr int V = 16 / (int)sizeof(T);
          using Vec = typename std::conditional<sizeof(T) == 4, int4, longlong2>::type;
          union { Vec q[CO_T / V]; T e[CO_T]; } wu;      // one broadcast 128-bit load per V channels
#pragma unroll
          for (int g = 0; g < CO_T / V; g++) wu.q[g] = reinterpret_cast<const Vec*>(wk)[g];
          const T* wv = wu.e;
          wk += a.CO_B;
#pragma unroll
          for (int j = 0; j < PX_T; j++) {
            const T xv = in_s[boff[j] + koff];
#pragma unroll
            for (int c = 0; c < CO_T; c++) acc[c][j] = mac<T>(wv[c], xv, acc[c][j]);
          }
        }
      }
    }
    // ---- epilogue: + bias, NCHW stores (consecutive pixel groups -> consecutive addresses)
    const int64_t chan_pix = (int64_t)a.HO * a.WO;
    T* ybase = a.y + (n0 * a.CO * (int64_t)a.HO + ho0) * a.WO;
#pragma unroll
    for (int c = 0; c < CO_T; c++) {
      const int co = co0 + cg * CO_T + c;
      if (co < a.CO) {
        const T b = a.bias ? a.bias[co] : T(0);
        T* yc = ybase + co * chan_pix;
#pragma unroll
        for (int j = 0; j < PX_T; j++)
          if (ooff[j] >= 0) yc[ooff[j]] = add_nocontract<T>(acc[c][j], b);
      }
    }
  }
}

// Picks a tiling; returns false when the shape does not fit (caller falls back to the gather kernel).
template <class T>
static bool plan_direct(DirectArgs<T>& a, int* co_t, int* nthreads, size_t* smem, int* grid_y) {
  const int KP = a.C * a.kH * a.kW;
  a.KP = KP;
  const size_t budget = 200 * 1024;
  // channel tile: 8 unless padding CO to a multiple of 8 wastes > 15 %
  auto padded = [](int v, int m) { return (v + m - 1) / m * m; };
  int CT = 8;
  if ((double)padded(a.CO, 8) / a.CO > 1.15 || sizeof(T) == 8) CT = 4;   // 8-byte accumulators: 4x8 tile fits 128 registers
  // channels per CTA: all of them if the weight slice stays under half of the budget
  int CO_B = padded(a.CO, CT);
  const size_t w_budget = budget * 3 / 5;
  while ((size_t)KP * CO_B * sizeof(T) > w_budget && CO_B > CT) CO_B = padded((CO_B + 1) / 2, CT);
  if ((size_t)KP * CO_B * sizeof(T) > w_budget) return false;
  const int CG = CO_B / CT;
  if (CG > 32) return false;
  int PG = 32 * (512 / (32 * CG));
  if (PG < 32) PG = 32;
  if (PG > 256) PG = 256;
  if (CG * PG > 512) return false;
  // input region: whole output height if it fits, otherwise a band of rows
  const int IW_T = (a.WO - 1) * a.sW + (a.kW - 1) * a.dW + 1;
  int IW_S = IW_T | 1;                                   // odd pitch: rows land in different banks
  const size_t in_budget = budget - (size_t)KP * CO_B * sizeof(T);
  int TH = a.HO;
  auto in_bytes = [&](int th, int imgs) {
    const int ih = (th - 1) * a.sH + (a.kH - 1) * a.dH + 1;
    return (size_t)imgs * a.C * ih * IW_S * sizeof(T);
  };
  while (TH > 1 && in_bytes(TH, 1) > in_budget) TH = (TH + 1) / 2;
  if (in_bytes(TH, 1) > in_budget) return false;
  int IMGS = 1;
  if (TH == a.HO) {
    // several images per CTA: amortise the weight staging, but keep >= ~2 CTAs per SM in flight
    const int64_t max_by_grid = a.N / (2 * (int64_t)sm_count()) > 0 ? a.N / (2 * (int64_t)sm_count()) : 1;
    while (IMGS + 1 <= max_by_grid && IMGS < 32 && in_bytes(TH, IMGS + 1) <= in_budget) IMGS++;
  }
  a.IMGS = IMGS; a.TH = TH; a.IW_T = IW_T; a.IW_S = IW_S;
  a.IH_T = (TH - 1) * a.sH + (a.kH - 1) * a.dH + 1;
  a.CO_B = CO_B; a.CG = CG; a.PG = PG;
  a.bands = (a.HO + TH - 1) / TH;
  *co_t = CT;
  *nthreads = CG * PG;
  *smem = (size_t)KP * CO_B * sizeof(T) + in_bytes(TH, IMGS);
  *grid_y = (a.CO + CO_B - 1) / CO_B;
  if ((int64_t)a.IMGS * a.C * a.IH_T * a.IW_S >= (1ll << 30)) return false;
  if ((int64_t)a.IMGS * a.CO * a.HO * a.WO >= (1ll << 31)) return false;     // 32-bit output offsets inside a CTA
  return true;
}

template <class T>
static int launch_direct(cudaStream_t st, DirectArgs<T>& a, bool* done) {
  *done = false;
  int co_t = 0, nthreads = 0, grid_y = 0;
  size_t smem = 0;
  if (!plan_direct<T>(a, &co_t, &nthreads, &smem, &grid_y)) return AM_OK;
  const int64_t gx = ceil_div(a.N, a.IMGS) * a.bands;
  if (gx > 2147483647ll || grid_y > 65535) return AM_OK;
  auto k4 = conv_direct_kernel<T, 4, 8>;
  auto kern = k4;
  if constexpr (sizeof(T) == 4) {
    if (co_t == 8) kern = conv_direct_kernel<T, 8, 8>;
  }
  AM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3((unsigned)gx, (unsigned)grid_y), nthreads, smem, st>>>(a);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  *done = true;
  return AM_OK;
}

// forward: y = conv(x, w) + bias.  *done = false -> caller must use the fallback.
template <class T>
int conv2d_forward_direct(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const T* input,
                          const T* kernel, const T* bias, T* output, bool* done) {
  DirectArgs<T> a{};
  a.x = input; a.w = kernel; a.bias = bias; a.y = output; a.N = d.N;
  a.C = (int)d.C; a.H = (int)d.H; a.W = (int)d.W; a.CO = (int)d.Cout; a.kH = (int)d.kH; a.kW = (int)d.kW;
  a.padH = (int)d.padH; a.padW = (int)d.padW; a.sH = (int)d.strideH; a.sW = (int)d.strideW;
  a.dH = (int)d.dilH; a.dW = (int)d.dilW; a.HO = (int)Ho; a.WO = (int)Wo;
  a.w_off = 0; a.w_sco = d.C * d.kH * d.kW; a.w_sci = d.kH * d.kW; a.w_skh = d.kW; a.w_skw = 1;
  return launch_direct<T>(st, a, done);
}

// data gradient for stride 1: grad_input = conv(grad_output, flip(W)^T) with padding d*(k-1) - p.
template <class T>
int conv2d_dgrad_direct(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const T* grad_output,
                        const T* kernel, T* grad_input, bool* done) {
  *done = false;
  if (d.strideH != 1 || d.strideW != 1) return AM_OK;
  DirectArgs<T> a{};
  a.x = grad_output; a.w = kernel; a.bias = nullptr; a.y = grad_input; a.N = d.N;
  a.C = (int)d.Cout; a.H = (int)Ho; a.W = (int)Wo; a.CO = (int)d.C; a.kH = (int)d.kH; a.kW = (int)d.kW;
  a.padH = (int)(d.dilH * (d.kH - 1) - d.padH); a.padW = (int)(d.dilW * (d.kW - 1) - d.padW);
  a.sH = 1; a.sW = 1; a.dH = (int)d.dilH; a.dW = (int)d.dilW; a.HO = (int)d.H; a.WO = (int)d.W;
  // "output channel" = ci of the original weights, "input channel" = co, taps flipped
  a.w_sco = d.kH * d.kW; a.w_sci = d.C * d.kH * d.kW; a.w_skh = -d.kW; a.w_skw = -1;
  a.w_off = (d.kH - 1) * d.kW + (d.kW - 1);
  return launch_direct<T>(st, a, done);
}

#define INST(T)                                                                                                  \
  template int conv2d_forward_direct<T>(cudaStream_t, const am_conv2d_desc&, int64_t, int64_t, const T*,         \
                                        const T*, const T*, T*, bool*);                                          \
  template int conv2d_dgrad_direct<T>(cudaStream_t, const am_conv2d_desc&, int64_t, int64_t, const T*, const T*, \
                                      T*, bool*);
INST(float)
INST(double)
INST(int32_t)
INST(int64_t)
#undef INST

}  // namespace am
