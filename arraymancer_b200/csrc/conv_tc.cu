// float32 conv2d forward / stride-1 data gradient as a fused implicit GEMM on the 5th-gen tensor cores
// (tcgen05 kind::tf32, 3xTF32 split, fp32 accuracy).  The im2col matrix never exists in HBM *or* as a whole in
// shared memory: per 32-wide k block the gather warps read the needed NCHW input pixels (L1/L2-cached, lanes along
// consecutive output pixels), split them into tf32 hi/lo and write the 128-pixel x 32-k operand tile straight into
// the 128B-swizzled K-major layout the UMMA descriptor expects; the weights (small) come pre-split via TMA.
//
//   D[p, co] = sum_k im2col[p, k] * W[co, k],  M = 128 output pixels per tile (TMEM lanes), N = Cout (<= 64 per CTA),
//   K = C*kH*kW.  GEMM view of conv.nim:81-106 (forward) and of col2im(W^T gout) (conv.nim:136-139, gather form).
//
// Persistent, warp-specialised CTA (384 threads), tiles strided over the grid:
//   warp 0      TMA producer for the weight tiles (hi / lo planes, K-major, packed by conv_tc_pack_weights_kernel)
//   warp 1      UMMA issuer: per 8-wide k step three tcgen05.mma (lo*hi, hi*lo, hi*hi), accumulation chains of
//               `flush_kb` k blocks into one of two TMEM buffers (the tensor core truncates when it accumulates,
//               see gemm_f32_tc.cu)
//   warp 2      TMEM allocation
//   warps 4-7   gather / split / swizzled-store of the im2col operand tile (thread r <-> output pixel r of the tile)
//   warps 8-11  drain finished chains (tcgen05.ld) into register accumulators with round-to-nearest adds, then the
//               epilogue: + bias, NCHW stores (lanes = consecutive pixels: coalesced)
#include <cuda.h>
#include <cstdlib>

#include "am_common.cuh"
#include "gemm_dispatch.h"
#include "ptx_sm100.cuh"

namespace am {

struct ConvTcArgs {
  const float* x;       // [N][C][H][W]        (grad_output for dgrad)
  const float* bias;    // [CO] or null
  float* y;             // [N][CO][HO][WO]     (grad_input for dgrad)
  const int2* tab;      // k -> {offset ci*H*W + kh*dH*W + kw*dW, (kh*dH << 16) | (kw*dW)}; Kpad entries
  int64_t P;            // total output pixels N*HO*WO
  int C, H, W, CO, HO, WO, padH, padW, sH, sW;
  int K, kblocks;       // K' = C*kH*kW, ceil(K'/32)
  int NP;               // Cout padded to a multiple of 16 (<= 64): UMMA N and weight-tile rows
  int flush_kb;
  int ntiles;
};

constexpr int kCtStages = 4;
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

__global__ void conv_tc_table_kernel(int2* tab, int K, int Kpad, int kH, int kW, int H, int W, int dH, int dW) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < Kpad; k += gridDim.x * blockDim.x) {
    if (k < K) {
      const int ci = k / (kH * kW), r = k - ci * (kH * kW), kh = r / kW, kw = r - kh * kW;
      tab[k] = make_int2(ci * H * W + kh * dH * W + kw * dW, ((kh * dH) << 16) | (kw * dW));
    } else {
      tab[k] = make_int2(0, 0x7fff7fff);      // out-of-range taps: h0 + 32767 is never inside the image
    }
  }
}

__global__ void __launch_bounds__(384, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo, const ConvTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)a.NP * 128u;                     // one weight plane per stage
  const uint32_t stage_bytes = 2u * kCtABytes + 2u * b_bytes;
  auto stage_base = [&](int s) { return smem_base + (uint32_t)s * stage_bytes; };
  const uint32_t OFF_AHI = 0, OFF_ALO = kCtABytes, OFF_BHI = 2 * kCtABytes, OFF_BLO = 2 * kCtABytes + b_bytes;
  const uint32_t bar_base = smem_base + kCtStages * stage_bytes;
  auto full_a = [&](int s) { return bar_base + 8u * s; };
  auto full_b = [&](int s) { return bar_base + 8u * (kCtStages + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * kCtStages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (3 * kCtStages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (3 * kCtStages + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * kCtStages + 4);
  const uint32_t tab_s = bar_base + 8u * (3 * kCtStages + 6);           // int2 tab[kblocks*32]

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = ptx::lane_id();
  const int nkb = a.kblocks;
  const uint32_t tmem_cols = (2u * (uint32_t)a.NP <= 32u) ? 32u : ((2u * a.NP <= 64u) ? 64u : 128u);

  if (warp == 0 && ptx::elect_one()) { ptx::prefetch_tensormap(&tmWhi); ptx::prefetch_tensormap(&tmWlo); }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < kCtStages; s++) {
      ptx::mbar_init(full_a(s), 128);      // every gather thread arrives after its swizzled stores
      ptx::mbar_init(full_b(s), 1);        // TMA transaction bytes
      ptx::mbar_init(empty_bar(s), 1);     // tcgen05.commit
    }
    for (int b = 0; b < 2; b++) { ptx::mbar_init(tfull_bar(b), 1); ptx::mbar_init(tempty_bar(b), 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<1>(tmem_slot, tmem_cols);
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

  if (warp == 0) {
    // ===================== TMA producer: weight tiles =====================
    if (ptx::elect_one()) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % kCtStages;
          ptx::mbar_wait(empty_bar(s), ((it / kCtStages) & 1u) ^ 1u);
          const uint32_t sb = stage_base(s);
          ptx::mbar_arrive_expect_tx(full_b(s), 2 * b_bytes);
          ptx::tma_load_2d(sb + OFF_BHI, &tmWhi, full_b(s), kb * 32, 0);
          ptx::tma_load_2d(sb + OFF_BLO, &tmWlo, full_b(s), kb * 32, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    if (ptx::elect_one()) {
      const uint64_t dhi = ptx::umma_desc_hi(1024, 128);
      const uint32_t idesc = ptx::umma_idesc_tf32(128, (uint32_t)a.NP);
      uint32_t it = 0, chain = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        int kb = 0;
        for (int c = 0; c < chains_per_tile; c++, chain++) {
          const int buf = chain & 1;
          ptx::mbar_wait(tempty_bar(buf), ((chain >> 1) & 1u) ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)buf * (uint32_t)a.NP;
          const int kb_end = (kb + flush < nkb) ? kb + flush : nkb;
          for (bool first = true; kb < kb_end; kb++, it++) {
            const int s = it % kCtStages;
            const uint32_t ph = (it / kCtStages) & 1u;
            ptx::mbar_wait(full_a(s), ph);
            ptx::mbar_wait(full_b(s), ph);
            ptx::tc_fence_after();
            const uint32_t sb = stage_base(s);
#pragma unroll
            for (int k8 = 0; k8 < 4; k8++) {
              const uint32_t koff = k8 * 32;
              const uint64_t a_hi = ptx::umma_desc(dhi, sb + OFF_AHI + koff), a_lo = ptx::umma_desc(dhi, sb + OFF_ALO + koff);
              const uint64_t b_hi = ptx::umma_desc(dhi, sb + OFF_BHI + koff), b_lo = ptx::umma_desc(dhi, sb + OFF_BLO + koff);
              ptx::umma_tf32<1>(d, a_lo, b_hi, idesc, first ? 0u : 1u);
              ptx::umma_tf32<1>(d, a_hi, b_lo, idesc, 1u);
              ptx::umma_tf32<1>(d, a_hi, b_hi, idesc, 1u);
              first = false;
            }
            ptx::umma_commit<1>(empty_bar(s));
          }
          ptx::umma_commit<1>(tfull_bar(buf));
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== gather warps: build the im2col operand tile =====================
    const int r = threadIdx.x - 128;                       // pixel row of the tile
    const uint32_t row_off = (uint32_t)r * 128u;
    const uint32_t sw = (uint32_t)(r & 7);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      const int64_t p = (int64_t)tile * 128 + r;
      const bool pok = p < a.P;
      const int64_t pp = pok ? p : 0;
      const int64_t n = pp / ((int64_t)a.HO * a.WO);
      const int rem = (int)(pp - n * (int64_t)a.HO * a.WO);
      const int ho = rem / a.WO, wo = rem - ho * a.WO;
      const int h0 = pok ? ho * a.sH - a.padH : -100000, w0 = wo * a.sW - a.padW;
      const float* xb = a.x + n * (int64_t)a.C * a.H * a.W + (int64_t)h0 * a.W + w0;
      for (int kb = 0; kb < nkb; kb++, it++) {
        const int s = it % kCtStages;
        ptx::mbar_wait(empty_bar(s), ((it / kCtStages) & 1u) ^ 1u);
        const uint32_t sb = stage_base(s);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) {
          int ex, ey;
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(ex), "=r"(ey) : "r"(tab_s + 8u * (uint32_t)(kb * 32 + j)));
          const int h = h0 + (ey >> 16), w = w0 + (ey & 0xffff);
          v[j] = ((unsigned)h < (unsigned)a.H && (unsigned)w < (unsigned)a.W) ? __ldg(xb + ex) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 8; c++) {
          float hi4[4], lo4[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const float xv = v[c * 4 + e];
            float h = xv, l = 0.f;
            if (isfinite(xv)) { h = ptx::to_tf32_rna(xv); l = ptx::to_tf32_rna(xv - h); }
            hi4[e] = h; lo4[e] = l;
          }
          const uint32_t off = row_off + (((uint32_t)c ^ sw) << 4);     // 128B swizzle: 16-byte chunk c of row r lands at c ^ (r & 7)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb + OFF_AHI + off), "f"(hi4[0]), "f"(hi4[1]), "f"(hi4[2]), "f"(hi4[3]) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb + OFF_ALO + off), "f"(lo4[0]), "f"(lo4[1]), "f"(lo4[2]), "f"(lo4[3]) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy stores -> visible to the UMMA (async proxy)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_a(s)) : "memory");
      }
    }
  } else if (warp >= 8) {
    // ===================== accumulate / epilogue warps =====================
    const int q = warp & 3;
    const int r = q * 32 + (int)lane;                      // TMEM lane = pixel row of the tile
    uint32_t chain = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      float acc[64];
#pragma unroll
      for (int i = 0; i < 64; i++) acc[i] = 0.f;
      for (int c = 0; c < chains_per_tile; c++, chain++) {
        const int buf = chain & 1;
        ptx::mbar_wait(tfull_bar(buf), (chain >> 1) & 1u);
        ptx::tc_fence_after();
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * (uint32_t)a.NP;
        if (a.NP > 32) {
          uint32_t r0[32], r1[32];
          ptx::tmem_ld_32x32(t0, r0);
          ptx::tmem_ld_32x32(t0 + 32, r1);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) { acc[i] = __fadd_rn(acc[i], __uint_as_float(r0[i])); acc[32 + i] = __fadd_rn(acc[32 + i], __uint_as_float(r1[i])); }
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
      // epilogue: out[n][co][ho][wo] = acc[co] + bias[co]; consecutive lanes = consecutive pixels
      const int64_t p = (int64_t)tile * 128 + r;
      if (p < a.P) {
        const int64_t hw = (int64_t)a.HO * a.WO;
        const int64_t n = p / hw, rem = p - n * hw;
        float* yp = a.y + n * a.CO * hw + rem;
#pragma unroll
        for (int co = 0; co < 64; co++)
          if (co < a.CO) yp[co * hw] = __fadd_rn(acc[co], a.bias ? a.bias[co] : 0.f);
      }
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
};

static int run_conv_tc(cudaStream_t st, const ConvTcView& v, bool* done) {
  *done = false;
  if (!gemm_f32_tc_available()) return AM_OK;
  const int K = v.C * v.kH * v.kW;
  if (v.CO > 64 || K > 4096) return AM_OK;                       // accumulators: 64 registers per pixel; table in smem
  if (v.kH * v.dH >= 32767 || v.kW * v.dW >= 32767) return AM_OK;
  if ((int64_t)v.C * v.H * v.W >= (1ll << 31)) return AM_OK;
  const int64_t P = v.N * (int64_t)v.HO * v.WO;
  if (P <= 0 || P >= (1ll << 37)) return AM_OK;
  const int Kpad = (int)round_up(K, 32), nkb = Kpad / 32;
  const int NP = (int)round_up(v.CO, 16);
  const int Rpad = 64;                                            // weight planes padded to 64 rows
  // workspace: weight hi/lo planes + k table
  void* ws = nullptr;
  const size_t plane = (size_t)Rpad * Kpad * sizeof(float);
  int rc = workspace(kWsConvW, 2 * plane + (size_t)Kpad * sizeof(int2) + 256, &ws);
  if (rc) return rc;
  float* whi = (float*)ws; float* wlo = whi + (size_t)Rpad * Kpad;
  int2* tab = (int2*)((char*)ws + 2 * plane);
  conv_tc_pack_weights_kernel<<<(unsigned)ceil_div((int64_t)Rpad * Kpad, 256), 256, 0, st>>>(
      v.w, whi, wlo, v.CO, Rpad, K, Kpad, v.C, v.kH, v.kW, v.w_off, v.w_sco, v.w_sci, v.w_skh, v.w_skw);
  conv_tc_table_kernel<<<(unsigned)ceil_div(Kpad, 256), 256, 0, st>>>(tab, K, Kpad, v.kH, v.kW, v.H, v.W, v.dH, v.dW);
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
  a.x = v.x; a.bias = v.bias; a.y = v.y; a.tab = tab; a.P = P;
  a.C = v.C; a.H = v.H; a.W = v.W; a.CO = v.CO; a.HO = v.HO; a.WO = v.WO; a.padH = v.padH; a.padW = v.padW; a.sH = v.sH; a.sW = v.sW;
  a.K = K; a.kblocks = nkb; a.NP = NP;
  static int flush_env = -1;
  if (flush_env < 0) { const char* e = getenv("AM_TC_FLUSH_KB"); flush_env = (e && atoi(e) > 0) ? atoi(e) : 2; }
  a.flush_kb = flush_env;
  a.ntiles = (int)ceil_div(P, 128);
  const size_t stage_bytes = 2 * (size_t)kCtABytes + 2 * (size_t)NP * 128;
  const size_t smem = kCtStages * stage_bytes + 1024 + 8 * (3 * kCtStages + 6) + (size_t)Kpad * 8 + 64;
  if (smem > 227 * 1024) return AM_OK;
  AM_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = a.ntiles < sm_count() ? a.ntiles : sm_count();
  conv_tc_kernel<<<grid, 384, smem, st>>>(tms[0], tms[1], a);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  *done = true;
  return AM_OK;
}

int conv2d_forward_tc_f32(cudaStream_t st, const am_conv2d_desc& d, int64_t Ho, int64_t Wo, const float* input,
                          const float* kernel, const float* bias, float* output, bool* done) {
  ConvTcView v{};
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
