// Implicit-GEMM conv2d forward / backward: the im2col buffer of the reference
// (nn_primitives/fallback/conv.nim:18-54, written then re-read by GEMM, 5-10x the algorithmic
// bytes) is never materialised — operand tiles are gathered straight from the NCHW tensors
// into shared memory by the loaders below and consumed by the contraction core.
//
//   forward  (conv.nim:81-106)   out[n,co,p]   = bias[co] + sum_k W[co,k] * im2col(in[n])[k,p]
//       GEMM view: M = Cout, N = Nimg*Ho*Wo, K = C*kH*kW           (batched over images, one launch)
//   dgrad    (conv.nim:136-139)  gin[n,ci,q]   = sum_{co,kh,kw} W[co,ci,kh,kw] * gout[n,co,ho,wo]
//       GEMM view: M = C, N = Nimg*H*W, K = Cout*kH*kW             (gather form of col2im: no atomics)
//   wgrad    (conv.nim:140)      gW[co,k]      = sum_{n,p} gout[n,co,p] * im2col(in[n])[k,p]
//       GEMM view: M = Cout, N = C*kH*kW (+1 ones-column = grad_bias, nnp_convolution.nim:94),
//                  K = Nimg*Ho*Wo split over CTAs, fixed-order second pass (deterministic).
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "contract_simt.cuh"
#include "gemm_dispatch.h"

namespace am {

struct ConvGeom {
  int64_t Nimg, C, H, W, Cout, kH, kW, padH, padW, sH, sW, dH, dW, Ho, Wo;
  int64_t HoWo, HW, kHkW, Kc;
};

static bool make_geom(const am_conv2d_desc& d, ConvGeom* g) {
  if (d.N < 0 || d.C <= 0 || d.H <= 0 || d.W <= 0 || d.Cout <= 0 || d.kH <= 0 || d.kW <= 0) return false;
  if (d.padH < 0 || d.padW < 0 || d.strideH <= 0 || d.strideW <= 0 || d.dilH <= 0 || d.dilW <= 0) return false;
  const int64_t eh = d.dilH * (d.kH - 1) + 1, ew = d.dilW * (d.kW - 1) + 1;
  if (d.H + 2 * d.padH < eh || d.W + 2 * d.padW < ew) return false;
  g->Nimg = d.N; g->C = d.C; g->H = d.H; g->W = d.W; g->Cout = d.Cout; g->kH = d.kH; g->kW = d.kW;
  g->padH = d.padH; g->padW = d.padW; g->sH = d.strideH; g->sW = d.strideW; g->dH = d.dilH; g->dW = d.dilW;
  g->Ho = (d.H + 2 * d.padH - eh) / d.strideH + 1;
  g->Wo = (d.W + 2 * d.padW - ew) / d.strideW + 1;
  g->HoWo = g->Ho * g->Wo; g->HW = d.H * d.W; g->kHkW = d.kH * d.kW; g->Kc = d.C * g->kHkW;
  // the gather tables use 32-bit offsets within one image / the weight tensor
  if (d.C * g->HW >= (1ll << 31) || d.Cout * g->HoWo >= (1ll << 31) || d.Cout * g->Kc >= (1ll << 31)) return false;
  if (d.dilH * d.kH >= 32768 || d.dilW * d.kW >= 32768) return false;
  return true;
}

// k-decode tables (int2 per k): x = element offset contribution, y = (kh*dH << 16) | (kw*dW)
//   tabF[k], k = (ci,kh,kw) : x = ci*H*W + kh*dH*W + kw*dW                  (fwd / wgrad input gather)
//   tabD[k], k = (co,kh,kw) : x = co*Ho*Wo                                  (dgrad grad_output gather)
//   tabW[k], k = (co,kh,kw) : x = co*C*kH*kW + kh*kW + kw                   (dgrad weight gather)
__global__ void conv_build_tables(ConvGeom g, int2* tabF, int2* tabD, int2* tabW) {
  const int64_t KD = g.Cout * g.kHkW;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < g.Kc || k < KD;
       k += (int64_t)gridDim.x * blockDim.x) {
    const int kw = (int)(k % g.kW), kh = (int)((k / g.kW) % g.kH);
    const int64_t c = k / g.kHkW;   // ci for tabF, co for tabD/tabW
    const int packed = (int)((kh * g.dH) << 16) | (int)(kw * g.dW);
    if (k < g.Kc) tabF[k] = make_int2((int)(c * g.HW + kh * g.dH * g.W + kw * g.dW), packed);
    if (k < KD) {
      tabD[k] = make_int2((int)(c * g.HoWo), packed);
      tabW[k] = make_int2((int)(c * g.C * g.kHkW + kh * g.kW + kw), packed);
    }
  }
}

// ------------------------------------------------------------------ loaders
// B operand of forward: column j = (n, ho, wo), row k = (ci, kh, kw) -> in[n, ci, ho*sH-padH+kh*dH, ...]
template <class T>
struct Im2colColsLoader {
  static constexpr int kMapping = 1;   // lanes along output pixels (contiguous in wo)
  const T* in; const int2* tab; ConvGeom g; int64_t NP;
  struct Slot { const T* base; int h0, w0, k; bool ok; };
  __device__ __forceinline__ Slot slot(int64_t j, int k) const {
    Slot s; s.k = k; s.ok = j < NP;
    const int64_t jj = s.ok ? j : 0;
    const int64_t n = jj / g.HoWo; const int p = (int)(jj - n * g.HoWo);
    const int ho = p / (int)g.Wo, wo = p - ho * (int)g.Wo;
    s.h0 = ho * (int)g.sH - (int)g.padH; s.w0 = wo * (int)g.sW - (int)g.padW;
    s.base = in + n * g.C * g.HW + (int64_t)s.h0 * g.W + s.w0;
    return s;
  }
  __device__ __forceinline__ T load(const Slot& s, int64_t kt) const {
    const int64_t k = kt + s.k;
    if (!s.ok || k >= g.Kc) return T(0);
    const int2 e = __ldg(tab + k);
    const int h = s.h0 + (e.y >> 16), w = s.w0 + (e.y & 0xffff);
    if ((unsigned)h >= (unsigned)g.H || (unsigned)w >= (unsigned)g.W) return T(0);
    return s.base[e.x];
  }
};

// A operand of dgrad: row m = ci, k = (co, kh, kw) -> W[co, ci, kh, kw]
template <class T>
struct WeightTLoader {
  static constexpr int kMapping = 2;
  const T* w; const int2* tab; ConvGeom g; int64_t KD;
  struct Slot { const T* base; int k; bool ok; };
  __device__ __forceinline__ Slot slot(int64_t ci, int k) const {
    return Slot{w + ci * g.kHkW, k, ci < g.C};
  }
  __device__ __forceinline__ T load(const Slot& s, int64_t kt) const {
    const int64_t k = kt + s.k;
    if (!s.ok || k >= KD) return T(0);
    return s.base[__ldg(tab + k).x];
  }
};

// B operand of dgrad: column j = (n, h, w), k = (co, kh, kw) -> gout[n, co, (h+padH-kh*dH)/sH, ...]
template <class T>
struct GradOutColsLoader {
  static constexpr int kMapping = 1;
  const T* gout; const int2* tab; ConvGeom g; int64_t NQ, KD;
  struct Slot { const T* base; int hp, wp, k; bool ok; };
  __device__ __forceinline__ Slot slot(int64_t j, int k) const {
    Slot s; s.k = k; s.ok = j < NQ;
    const int64_t jj = s.ok ? j : 0;
    const int64_t n = jj / g.HW; const int q = (int)(jj - n * g.HW);
    const int h = q / (int)g.W, w = q - h * (int)g.W;
    s.hp = h + (int)g.padH; s.wp = w + (int)g.padW;
    s.base = gout + n * g.Cout * g.HoWo;
    return s;
  }
  __device__ __forceinline__ T load(const Slot& s, int64_t kt) const {
    const int64_t k = kt + s.k;
    if (!s.ok || k >= KD) return T(0);
    const int2 e = __ldg(tab + k);
    int hh = s.hp - (e.y >> 16), ww = s.wp - (e.y & 0xffff);
    if (hh < 0 || ww < 0) return T(0);
    if (g.sH != 1) { if (hh % (int)g.sH) return T(0); hh /= (int)g.sH; }
    if (g.sW != 1) { if (ww % (int)g.sW) return T(0); ww /= (int)g.sW; }
    if (hh >= (int)g.Ho || ww >= (int)g.Wo) return T(0);
    return s.base[e.x + hh * (int)g.Wo + ww];
  }
};

// A operand of wgrad: row m = co, k = (n, p) -> gout[n, co, p]
template <class T>
struct GradOutRowsLoader {
  static constexpr int kMapping = 2;   // lanes along k = pixels (contiguous)
  const T* gout; ConvGeom g; int64_t NP;
  struct Slot { const T* base; int k; bool ok; };
  __device__ __forceinline__ Slot slot(int64_t co, int k) const {
    return Slot{gout + co * g.HoWo, k, co < g.Cout};
  }
  __device__ __forceinline__ T load(const Slot& s, int64_t kt) const {
    const int64_t k = kt + s.k;
    if (!s.ok || k >= NP) return T(0);
    const int64_t n = k / g.HoWo; const int64_t p = k - n * g.HoWo;
    return s.base[n * g.Cout * g.HoWo + p];
  }
};

// B operand of wgrad: column j = (ci, kh, kw) [j == Kc: the ones-column], k = (n, ho, wo)
template <class T>
struct Im2colRowsLoader {
  static constexpr int kMapping = 2;
  const T* in; ConvGeom g; int64_t NP; int with_bias_col;
  struct Slot { int64_t off; int khd, kwd, k, kind; };   // kind: 0 = zero, 1 = gather, 2 = ones
  __device__ __forceinline__ Slot slot(int64_t j, int k) const {
    Slot s; s.k = k; s.off = 0; s.khd = 0; s.kwd = 0;
    if (j < g.Kc) {
      const int kw = (int)(j % g.kW), kh = (int)((j / g.kW) % g.kH);
      const int64_t ci = j / g.kHkW;
      s.khd = kh * (int)g.dH; s.kwd = kw * (int)g.dW;
      s.off = ci * g.HW; s.kind = 1;
    } else {
      s.kind = (j == g.Kc && with_bias_col) ? 2 : 0;
    }
    return s;
  }
  __device__ __forceinline__ T load(const Slot& s, int64_t kt) const {
    const int64_t k = kt + s.k;
    if (s.kind == 0 || k >= NP) return T(0);
    if (s.kind == 2) return T(1);
    const int64_t n = k / g.HoWo; const int p = (int)(k - n * g.HoWo);
    const int ho = p / (int)g.Wo, wo = p - ho * (int)g.Wo;
    const int h = ho * (int)g.sH - (int)g.padH + s.khd, w = wo * (int)g.sW - (int)g.padW + s.kwd;
    if ((unsigned)h >= (unsigned)g.H || (unsigned)w >= (unsigned)g.W) return T(0);
    return in[n * g.C * g.HW + s.off + (int64_t)h * g.W + w];
  }
};

// ------------------------------------------------------------------ epilogues
// rows = channels, columns = (image, pixel): dst[(n*CH + ch)*PIX + p] = v (+ bias[ch])
template <class T>
struct NchwEpilogue {
  T* dst; const T* bias; int64_t CH, PIX, NCOLS; bool vec_ok; bool relu;
  __device__ __forceinline__ T fin(T v, T b) const {
    const T s = add_nocontract<T>(v, b);
    return (relu && s <= T(0)) ? T(0) : s;
  }
  template <int V>
  __device__ __forceinline__ void store(int64_t ch, int64_t j0, const T (&v)[V], int) const {
    if (ch >= CH || j0 >= NCOLS) return;
    const T b = bias ? bias[ch] : T(0);
    if (vec_ok && j0 + V <= NCOLS) {        // PIX % V == 0: the V pixels share an image and are contiguous
      const int64_t n = j0 / PIX, p = j0 - n * PIX;
      using Vec = typename std::conditional<sizeof(T) == 4, int4, longlong2>::type;
      union { Vec q; T e[V]; } o;
#pragma unroll
      for (int j = 0; j < V; j++) o.e[j] = fin(v[j], b);
      *reinterpret_cast<Vec*>(dst + (n * CH + ch) * PIX + p) = o.q;
    } else {
#pragma unroll
      for (int j = 0; j < V; j++) {
        const int64_t jj = j0 + j;
        if (jj < NCOLS) {
          const int64_t n = jj / PIX, p = jj - n * PIX;
          dst[(n * CH + ch) * PIX + p] = fin(v[j], b);
        }
      }
    }
  }
};

// split-K partials: part[z][m][n], m < Mv, n < Nv (dense)
template <class T>
struct PartialEpilogue {
  T* part; int64_t Mv, Nv;
  template <int V>
  __device__ __forceinline__ void store(int64_t m, int64_t n0, const T (&v)[V], int z) const {
    if (m >= Mv) return;
    T* row = part + ((int64_t)z * Mv + m) * Nv;
#pragma unroll
    for (int j = 0; j < V; j++)
      if (n0 + j < Nv) row[n0 + j] = v[j];
  }
};

// fixed-order reduction of the split-K partials -> grad_kernel [Cout, Kc] and grad_bias [Cout]
template <class T>
__global__ void wgrad_reduce_kernel(const T* part, int splits, int64_t Cout, int64_t Kc, int64_t Nv, T* gk, T* gb) {
  const int64_t total = Cout * Nv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    T acc = T(0);
    for (int z = 0; z < splits; z++) acc = add_nocontract<T>(acc, part[(int64_t)z * total + i]);
    const int64_t co = i / Nv, j = i - co * Nv;
    if (j < Kc) { if (gk) gk[co * Kc + j] = acc; }
    else if (gb) gb[co] = acc;
  }
}

// ------------------------------------------------------------------ table cache
struct TableCache {
  am_conv2d_desc desc{}; int device = -1; bool valid = false;
  int2 *tabF = nullptr, *tabD = nullptr, *tabW = nullptr;
  cudaStream_t stream = nullptr;     // the stream the tables were built on: a hit needs the same stream (stream order is the
                                     // only ordering used — no events, so the calls can be captured into CUDA graphs)
};
// One cache entry per DEVICE, shared by all host threads and guarded by a mutex: the tables live in the per-device
// workspace slot kWsConvTab, so the record of what that buffer holds must be per device too (a per-thread record let a
// second thread overwrite the buffer while the first thread still believed its descriptor was cached).
static TableCache g_tab[16];
static std::mutex g_tab_mu;

static int get_tables(cudaStream_t st, const am_conv2d_desc& d, const ConvGeom& g, int2** tabF, int2** tabD,
                      int2** tabW) {
  int dev = 0;
  AM_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) { set_last_error("conv: device index out of range"); return AM_ERR_INVALID; }
  const int64_t KD = g.Cout * g.kHkW;
  void* base = nullptr;
  const size_t nF = (size_t)round_up(g.Kc, 64), nD = (size_t)round_up(KD, 64);
  int rc = workspace(kWsConvTab, (nF + 2 * nD) * sizeof(int2), &base);
  if (rc) return rc;
  int2* f = (int2*)base; int2* dd = f + nF; int2* w = dd + nD;
  std::lock_guard<std::mutex> lk(g_tab_mu);
  TableCache& tc = g_tab[dev];
  // conv calls of one device run on one stream at a time (include/am_b200.h): a call on another stream rebuilds the tables
  // there (one tiny kernel) instead of waiting on an event of the first stream
  const bool hit = tc.valid && tc.tabF == f && tc.stream == st && memcmp(&tc.desc, &d, sizeof(d)) == 0;
  if (!hit) {
    const int64_t n = g.Kc > KD ? g.Kc : KD;
    conv_build_tables<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(g, f, dd, w);
    g_launch_count++;
    AM_CUDA_TRY(cudaGetLastError());
    tc.desc = d; tc.device = dev; tc.valid = true; tc.tabF = f; tc.tabD = dd; tc.tabW = w; tc.stream = st;
  }
  *tabF = f; *tabD = dd; *tabW = w;
  return AM_OK;
}
void conv_tables_invalidate() {
  std::lock_guard<std::mutex> lk(g_tab_mu);
  for (auto& t : g_tab) t.valid = false;
}

// ------------------------------------------------------------------ launchers
template <class T> struct ConvCfgs {
  static constexpr int BK = 64 / (int)sizeof(T);
  using C32 = SimtCfg<T, 32, 256, BK, 4, 8>;    // <= 32 channels on the M axis
  using C64 = SimtCfg<T, 64, 128, BK, 8, 4>;    // <= 64
  using C128 = SimtCfg<T, 128, 128, BK, 8, 8>;
  using W32 = SimtCfg<T, 32, 64, BK, 4, 4>;     // wgrad: small Cout x small Kc
};

template <class T, class Cfg, class LA, class LB, class Epi>
static int launch(cudaStream_t st, const LA& la, const LB& lb, const Epi& epi, int64_t M, int64_t N, int64_t K,
                  int splits, int64_t k_per_split) {
  dim3 grid((unsigned)ceil_div(N, Cfg::BN), (unsigned)ceil_div(M, Cfg::BM), (unsigned)splits);
  if (grid.y > 65535 || grid.z > 65535) { set_last_error("conv: grid too large"); return AM_ERR_INVALID; }
  contract_simt_kernel<T, Cfg, LA, LB, Epi><<<grid, Cfg::NT, 0, st>>>(la, lb, epi, K, k_per_split, 1, 0);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  return AM_OK;
}

template <class T, class LA, class LB, class Epi>
static int launch_by_rows(cudaStream_t st, const LA& la, const LB& lb, const Epi& epi, int64_t M, int64_t N,
                          int64_t K) {
  if (M <= 32) return launch<T, typename ConvCfgs<T>::C32>(st, la, lb, epi, M, N, K, 1, K);
  if (M <= 64) return launch<T, typename ConvCfgs<T>::C64>(st, la, lb, epi, M, N, K, 1, K);
  return launch<T, typename ConvCfgs<T>::C128>(st, la, lb, epi, M, N, K, 1, K);
}

// conv_direct.cu: float32 shared-memory-staged direct kernels (fast path); *done == false -> use the gather kernels
int conv2d_forward_direct_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* input,
                              const float* kernel, const float* bias, float* output, int act, bool* done);
int conv2d_dgrad_direct_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* grad_output,
                            const float* kernel, float* grad_input, bool* done);
int conv2d_wgrad_direct_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* input,
                            const float* grad_output, float** part_out, int* groups_out, bool* done);
// conv_tc.cu: tcgen05 implicit-GEMM kernels (float32, Cout <= 64); *done == false -> next path
int conv2d_forward_tc_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* input,
                          const float* kernel, const float* bias, float* output, int act, bool* done);
int conv2d_dgrad_tc_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* grad_output,
                        const float* kernel, float* grad_input, bool* done);
// conv_tc_bwd.cu: data gradient as GEMM + col2im on the tensor cores (any stride / padding / dilation, Cout <= 64)
int conv2d_dgrad_col2im_tc_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* grad_output,
                               const float* kernel, float* grad_input, bool only_if_fast, bool* done);
int conv2d_wgrad_tc_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* input,
                        const float* grad_output, float** part_out, int* groups_out, bool* done);
// conv_c1.cu: single-input-channel fused kernels (LeNet cv1 class); *done == false -> next path
int conv2d_forward_c1_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* input,
                          const float* kernel, const float* bias, float* output, int act, bool* done);
int conv2d_backward_c1_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* input,
                           const float* kernel, const float* grad_output, float* grad_input, float* grad_kernel,
                           float* grad_bias, bool* done);
std::atomic<int> g_conv_path{AM_CONV_AUTO};
static bool direct_enabled() { return g_conv_path.load() != AM_CONV_GATHER; }
static bool tc_enabled() { return g_conv_path.load() == AM_CONV_TC; }
static bool auto_path() { return g_conv_path.load() == AM_CONV_AUTO; }

template <class T>
int conv2d_forward(cudaStream_t st, const am_conv2d_desc& d, const T* input, const T* kernel, const T* bias,
                   T* output, int act) {
  ConvGeom g;
  if (!make_geom(d, &g)) { set_last_error("conv2d_forward: invalid geometry"); return AM_ERR_INVALID; }
  if (g.Nimg == 0) return AM_OK;
  if (!input || !kernel || !output) { set_last_error("conv2d_forward: null pointer"); return AM_ERR_INVALID; }
  if constexpr (std::is_same<T, float>::value) {
    if (direct_enabled() && !tc_enabled()) {          // C = 1: the fused HBM-bound kernel (AUTO and DIRECT)
      bool done = false;
      int rcd = conv2d_forward_c1_f32(st, d, g.Ho, g.Wo, input, kernel, bias, output, act, &done);
      if (rcd || done) return rcd;
    }
    // AUTO: the tcgen05 implicit GEMM pays once the GEMM is wide and deep enough (measured: LeNet cv2 yes, cv1 no)
    if (tc_enabled() || (auto_path() && d.Cout >= 32 && g.Kc >= 128)) {
      bool done = false;
      int rcd = conv2d_forward_tc_f32(st, d, g.Ho, g.Wo, input, kernel, bias, output, act, &done);
      if (rcd || done) return rcd;
    }
    if (direct_enabled()) {
      bool done = false;
      int rcd = conv2d_forward_direct_f32(st, d, g.Ho, g.Wo, input, kernel, bias, output, act, &done);
      if (rcd || done) return rcd;
    }
  }
  int2 *tabF, *tabD, *tabW;
  int rc = get_tables(st, d, g, &tabF, &tabD, &tabW);
  if (rc) return rc;
  const int64_t NP = g.Nimg * g.HoWo;
  StridedLoader<T> la{kernel, g.Kc, 1, g.Cout, g.Kc};
  Im2colColsLoader<T> lb{input, tabF, g, NP};
  constexpr int V = 16 / (int)sizeof(T);
  NchwEpilogue<T> epi{output, bias, g.Cout, g.HoWo, NP,
                      (g.HoWo % V == 0) && ((reinterpret_cast<uintptr_t>(output) & 15) == 0), act != 0};
  return launch_by_rows<T>(st, la, lb, epi, g.Cout, NP, g.Kc);
}

template <class T>
int conv2d_backward(cudaStream_t st, const am_conv2d_desc& d, const T* input, const T* kernel,
                    const T* grad_output, T* grad_input, T* grad_kernel, T* grad_bias) {
  ConvGeom g;
  if (!make_geom(d, &g)) { set_last_error("conv2d_backward: invalid geometry"); return AM_ERR_INVALID; }
  if (!grad_output || (grad_input && !kernel) || ((grad_kernel || grad_bias) && !input)) {
    // the weight-gradient kernels produce grad_bias as an extra column of the same pass, so they read `input` too
    set_last_error("conv2d_backward: null pointer (grad_kernel / grad_bias need `input`, grad_input needs `kernel`)");
    return AM_ERR_INVALID;
  }
  constexpr int V = 16 / (int)sizeof(T);
  const int64_t NP = g.Nimg * g.HoWo, NQ = g.Nimg * g.HW, KD = g.Cout * g.kHkW;
  int2 *tabF, *tabD, *tabW;
  int rc = get_tables(st, d, g, &tabF, &tabD, &tabW);
  if (rc) return rc;

  bool dgrad_done = false;
  if constexpr (std::is_same<T, float>::value) {
    if (g.Nimg > 0 && direct_enabled() && !tc_enabled() && (grad_input || grad_kernel || grad_bias)) {
      // C = 1: data, weight and bias gradients from ONE pass over grad_output
      bool done = false;
      rc = conv2d_backward_c1_f32(st, d, g.Ho, g.Wo, input, kernel, grad_output, grad_input, grad_kernel, grad_bias, &done);
      if (rc || done) return rc;
    }
    if (grad_input && g.Nimg > 0 && (tc_enabled() || auto_path())) {
      const bool gather_form = tuning(kTuneConvDgradGather) != 0;      // older gather-form kernel (comparison)
      if (gather_form && tc_enabled()) rc = conv2d_dgrad_tc_f32(st, d, g.Ho, g.Wo, grad_output, kernel, grad_input, &dgrad_done);
      else rc = conv2d_dgrad_col2im_tc_f32(st, d, g.Ho, g.Wo, grad_output, kernel, grad_input, !tc_enabled(), &dgrad_done);
      if (rc) return rc;
    }
    if (grad_input && g.Nimg > 0 && !dgrad_done && direct_enabled()) {
      rc = conv2d_dgrad_direct_f32(st, d, g.Ho, g.Wo, grad_output, kernel, grad_input, &dgrad_done);
      if (rc) return rc;
    }
  }
  if (grad_input && g.Nimg > 0 && !dgrad_done) {          // dgrad, gather form (any stride)
    WeightTLoader<T> la{kernel, tabW, g, KD};
    GradOutColsLoader<T> lb{grad_output, tabD, g, NQ, KD};
    NchwEpilogue<T> epi{grad_input, nullptr, g.C, g.HW, NQ,
                        (g.HW % V == 0) && ((reinterpret_cast<uintptr_t>(grad_input) & 15) == 0), false};
    rc = launch_by_rows<T>(st, la, lb, epi, g.C, NQ, KD);
    if (rc) return rc;
  }

  if constexpr (std::is_same<T, float>::value) {
    if ((grad_kernel || grad_bias) && direct_enabled() && g.Nimg > 0) {
      float* part = nullptr; int groups = 0; bool done = false;
      // tensor-core weight gradient: measured faster than the SIMT kernel on both LeNet layers (0.54 vs 0.64 ms, 0.27 vs 1.10 ms)
      if ((tc_enabled() || auto_path()) && input) {
        rc = conv2d_wgrad_tc_f32(st, d, g.Ho, g.Wo, input, grad_output, &part, &groups, &done);
        if (rc) return rc;
      }
      if (!done) rc = conv2d_wgrad_direct_f32(st, d, g.Ho, g.Wo, input, grad_output, &part, &groups, &done);
      if (rc) return rc;
      if (done) {
        const int64_t Nv = g.Kc + 1, total = g.Cout * Nv;
        wgrad_reduce_kernel<float><<<(unsigned)ceil_div(total, 128), 128, 0, st>>>(part, groups, g.Cout, g.Kc, Nv,
                                                                                  grad_kernel, grad_bias);
        g_launch_count++;
        AM_CUDA_TRY(cudaGetLastError());
        return AM_OK;
      }
    }
  }
  if (grad_kernel || grad_bias) {          // wgrad (+ grad_bias as a ones-column), split over the batch
    const int64_t Nv = g.Kc + 1;
    const bool small = g.Cout <= 32 && Nv <= 64;
    const int64_t bm = small ? 32 : (g.Cout <= 64 ? 64 : 128), bn = small ? 64 : 128;
    const int64_t tiles = ceil_div(g.Cout, bm) * ceil_div(Nv, bn);
    const int64_t bk = 64 / (int64_t)sizeof(T);
    int64_t splits = ceil_div(4 * (int64_t)sm_count(), tiles);
    const int64_t max_splits = NP > 0 ? ceil_div(NP, 8 * bk) : 1;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    const int64_t kps = NP > 0 ? round_up(ceil_div(NP, splits), bk) : bk;
    splits = NP > 0 ? ceil_div(NP, kps) : 1;
    void* part = nullptr;
    rc = workspace(kWsConv, (size_t)(splits * g.Cout * Nv) * sizeof(T), &part);
    if (rc) return rc;
    GradOutRowsLoader<T> la{grad_output, g, NP};
    Im2colRowsLoader<T> lb{input, g, NP, 1};
    PartialEpilogue<T> epi{(T*)part, g.Cout, Nv};
    if (small) rc = launch<T, typename ConvCfgs<T>::W32>(st, la, lb, epi, g.Cout, Nv, NP, (int)splits, kps);
    else if (g.Cout <= 64) rc = launch<T, typename ConvCfgs<T>::C64>(st, la, lb, epi, g.Cout, Nv, NP, (int)splits, kps);
    else rc = launch<T, typename ConvCfgs<T>::C128>(st, la, lb, epi, g.Cout, Nv, NP, (int)splits, kps);
    if (rc) return rc;
    const int64_t total = g.Cout * Nv;
    wgrad_reduce_kernel<T><<<(unsigned)ceil_div(total, 128), 128, 0, st>>>((const T*)part, (int)splits, g.Cout,
                                                                           g.Kc, Nv, grad_kernel, grad_bias);
    g_launch_count++;
    AM_CUDA_TRY(cudaGetLastError());
  }
  return AM_OK;
}

#define INST(T)                                                                                              \
  template int conv2d_forward<T>(cudaStream_t, const am_conv2d_desc&, const T*, const T*, const T*, T*, int); \
  template int conv2d_backward<T>(cudaStream_t, const am_conv2d_desc&, const T*, const T*, const T*, T*, T*, T*);
INST(float)
INST(double)
INST(int32_t)
INST(int64_t)
#undef INST

}  // namespace am
