// On-device micro-benchmarks that define the per-dtype roofline denominators (SURVEY §8d:
// no datasheet number is trusted).  Each one times a register-resident dependent-chain
// kernel (SIMT pipes) or a back-to-back tcgen05.mma issue loop (tensor pipe) with CUDA events.
#include "am_common.cuh"
#include "gemm_dispatch.h"
#include "ptx_sm100.cuh"

namespace am {

template <int CHAINS>
__global__ void __launch_bounds__(256) i64_narrow_peak_kernel(int64_t* out, int iters, int64_t a0, int64_t b0) {
  int64_t acc[CHAINS], bb[CHAINS];
  int64_t a = a0 + threadIdx.x;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) { acc[i] = i; bb[i] = b0 + 7 * i + (threadIdx.x & 3); }   // distinct per chain: products cannot be shared
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc[i] = mac_narrow_i64(a, bb[i], acc[i]);
    a = (int32_t)(a + acc[0]);
  }
  int64_t s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += acc[i];
  if (s == 123457) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class T, int CHAINS>
__global__ void __launch_bounds__(256) simt_peak_kernel(T* out, int iters, T a0, T b0) {
  T acc[CHAINS], bb[CHAINS];
  T a = a0 + (T)threadIdx.x;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) {
    acc[i] = (T)i;
    // a distinct multiplicand per chain, like the b-fragment of a register tile: integer products / cross terms
    // cannot be computed once and shared between chains (the first version of this benchmark allowed that and
    // over-stated the int64 rate: 15.6 TOP/s "peak" vs 9.3 from the IMAD-slot count)
    if constexpr (std::is_floating_point<T>::value) bb[i] = b0 + (T)(i * 0.001);
    else bb[i] = (T)(b0 + (T)(0x10001 * i) + (T)(threadIdx.x & 3) + (sizeof(T) == 8 ? ((T)i << 33) : (T)0));
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc[i] = mac<T>(a, bb[i], acc[i]);
    a = (T)(a + acc[0]);       // keep the multiplicand live and data dependent (1 extra op / CHAINS macs)
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s = (T)(s + acc[i]);
  if (s == (T)123457) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed FP32: fma.rn.f32x2 (SASS FFMA2) does two FMAs per lane and instruction on aligned 64-bit register pairs
template <int CHAINS>
__global__ void __launch_bounds__(256) ffma2_peak_kernel(float* out, int iters, float a0, float b0) {
  unsigned long long acc[CHAINS], bb[CHAINS], a;
  {
    const float ax = a0 + threadIdx.x * 1e-6f, ay = a0 - threadIdx.x * 1e-6f;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(ax), "f"(ay));
  }
#pragma unroll
  for (int i = 0; i < CHAINS; i++) {
    const float f = (float)i, g = b0 + i * 0.001f;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(f), "f"(-f));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(bb[i]) : "f"(g), "f"(g + 0.0005f));
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(a), "l"(bb[i]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) { float x, y; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i])); s += x + y; }
  if (s == 123457.f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mma.sync m8n8k4 f64: 8x8x4 = 256 FMA per warp instruction
template <int CHAINS>
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double c0[CHAINS], c1[CHAINS];
  double a = 1.0 + threadIdx.x * 1e-3, b = 0.999;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += c0[i] + c1[i];
  if (s == 123457.0) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// back-to-back tcgen05.mma kind::tf32 on resident operands: pure tensor-pipe rate.  Shape and operand placement are
// template parameters and the 16 MMAs of a batch are fully unrolled, so that the issuing thread executes nothing
// but UTCHMMAs between two commits (a first version of the small-N variants took N / accumulator / placement as
// run-time arguments: the per-MMA address arithmetic alone cut the measured N = 256 rate from 908 to 677 TFLOP/s).
//   N_MMA: UMMA N (M = 128 per CTA) | NACC: accumulators used round-robin | TS: A operand read from tensor memory
template <int CG, int N_MMA, int NACC, int TS>
__global__ void __launch_bounds__(128, 1) umma_peak_kernel(int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_s = base, b_s = base + 16384, bar = base + 16384 + 32768, slot = bar + 8;
  // pseudo-random tf32 operand bits (so the multipliers toggle like in a real GEMM)
  for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) {
    uint32_t x = i * 2654435761u + blockIdx.x * 40503u;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    const float f = ((int)(x & 0xffff) - 32768) * (1.0f / 32768.0f);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + 4 * i), "f"(ptx::to_tf32_rna(f)));
  }
  const int warp = threadIdx.x >> 5;
  const bool leader = (CG == 1) || ptx::cluster_ctarank() == 0;
  if (CG == 2) ptx::cluster_sync();
  if (warp == 0 && ptx::elect_one()) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
  if (warp == 1) ptx::tmem_alloc<CG>(slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (UMMA)
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (warp == 0 && leader && ptx::elect_one()) {
    const uint64_t dhi = ptx::umma_desc_hi(1024, 128);
    const uint32_t idesc = ptx::umma_idesc_tf32(128 * CG, (uint32_t)N_MMA);
    uint64_t adesc[4], bdesc[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { adesc[q] = ptx::umma_desc(dhi, a_s + q * 32); bdesc[q] = ptx::umma_desc(dhi, b_s + q * 32); }
    const uint32_t a_t = tmem + 480u;                          // TS: any 32 columns outside the accumulators
    uint32_t ph = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const uint32_t d = tmem + (uint32_t)((j % NACC) * N_MMA);          // NACC > 1: neighbours do not depend on each other
        const uint32_t en = (it > 0 || j >= NACC) ? 1u : 0u;
        if (TS) ptx::umma_tf32_ts(d, a_t + 8u * (j & 3), bdesc[j & 3], idesc, en);
        else ptx::umma_tf32<CG>(d, adesc[j & 3], bdesc[j & 3], idesc, en);
      }
      // single-CTA arrive even for CG == 2: only the leader waits
      asm volatile("tcgen05.commit.cta_group::%1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar), "n"(CG) : "memory");
      ptx::mbar_wait(bar, ph);
      ph ^= 1;
    }
  }
  __syncwarp();
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<CG>(tmem, 512);
}

// The issue pattern of the implicit-GEMM conv kernels, piece by piece (what does the single issuing thread pay per k block
// of 12 MMAs?).  PAT bit 0: three MMAs per k8 step on alternating A columns / B planes (3xTF32) instead of one operand pair;
// bit 1: two tcgen05.commit per k block (operand-stage releases); bit 2: two mbarrier waits (already complete) + fence per
// k block; bit 3: accumulation chains of 2 k blocks on alternating accumulators, one more commit per chain; bit 4: 16 KB
// per k block bulk-copied from a 256 KB L2-resident buffer into a 4-deep shared-memory ring (the weight stream of the
// conv kernels: every SM re-reads the same packed weights for every tile).
template <int PAT>
__global__ void __launch_bounds__(384, 1) umma_pattern_kernel(int iters, const float* __restrict__ gsrc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_hi_s = base, b_lo_s = base + 8192, bar = base + 16384, bar_e = bar + 8, bar_f = bar + 24, slot = bar + 40;
  const uint32_t bar_t = bar + 64, ring = base + 32768;      // bit 4: 4 transaction barriers + 4 x 16 KB landing ring
  for (uint32_t i = threadIdx.x; i < 16384 / 4; i += blockDim.x) {
    uint32_t x = i * 2654435761u + blockIdx.x * 40503u;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + 4 * i), "f"(ptx::to_tf32_rna(((int)(x & 0xffff) - 32768) * (1.0f / 32768.0f))));
  }
  const int warp = threadIdx.x >> 5;
  if (warp == 0 && ptx::elect_one()) {
    ptx::mbar_init(bar, 1); ptx::mbar_init(bar_e, 1); ptx::mbar_init(bar_e + 8, 1); ptx::mbar_init(bar_f, 1);
    for (int q = 0; q < 4; q++) ptx::mbar_init(bar_t + 8 * q, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<1>(slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (warp == 0 && ptx::elect_one()) {
    const uint64_t dhi = ptx::umma_desc_hi(1024, 128);
    const uint32_t idesc = ptx::umma_idesc_tf32(128, 64u);
    uint32_t ph = 0;
    for (int it = 0; it < iters; it++) {                      // one iteration = one k block of 32
      if (PAT & 16) {
        const int sq = it & 3;
        if (it >= 4) ptx::mbar_wait(bar_t + 8 * sq, (uint32_t)(((it >> 2) - 1) & 1));      // the copy issued 4 k blocks ago has landed
        ptx::mbar_arrive_expect_tx(bar_t + 8 * sq, 16384u);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(ring + 16384u * sq), "l"(gsrc + (size_t)(it & 15) * 4096), "r"(16384u), "r"(bar_t + 8 * sq) : "memory");
      }
      if (PAT & 4) {
        ptx::mbar_wait(bar_f, 1u);                            // fresh barrier, parity 1: returns at once
        ptx::mbar_wait(bar_f, 1u);
        ptx::tc_fence_after();
      }
      const uint32_t d = tmem + ((PAT & 8) ? (uint32_t)(((it >> 1) & 1) * 64) : 0u);
      const uint32_t a_hi0 = tmem + 128u + 64u * (uint32_t)(it % 6), a_lo0 = a_hi0 + 32u;
#pragma unroll
      for (int k8 = 0; k8 < 4; k8++) {
        const uint64_t bh = ptx::umma_desc(dhi, b_hi_s + k8 * 32), bl = ptx::umma_desc(dhi, b_lo_s + k8 * 32);
        const uint32_t first = ((PAT & 8) ? ((it & 1) == 0 && k8 == 0) : (it == 0 && k8 == 0)) ? 0u : 1u;
        if (PAT & 1) {
          ptx::umma_tf32_ts(d, a_lo0 + 8u * k8, bh, idesc, first);
          ptx::umma_tf32_ts(d, a_hi0 + 8u * k8, bl, idesc, 1u);
          ptx::umma_tf32_ts(d, a_hi0 + 8u * k8, bh, idesc, 1u);
        } else {
          ptx::umma_tf32_ts(d, a_hi0 + 8u * k8, bh, idesc, first);
          ptx::umma_tf32_ts(d, a_hi0 + 8u * k8, bh, idesc, 1u);
          ptx::umma_tf32_ts(d, a_hi0 + 8u * k8, bh, idesc, 1u);
        }
      }
      if (PAT & 2) { ptx::umma_commit<1>(bar_e); ptx::umma_commit<1>(bar_e + 8); }
      if ((PAT & 8) && (it & 1)) ptx::umma_commit<1>(bar_e);
      if ((it & 15) == 15) {                                  // bound the queue like the microbenchmark: wait every 192 MMAs
        ptx::umma_commit<1>(bar);
        ptx::mbar_wait(bar, ph);
        ph ^= 1;
      }
    }
  }
  // bit 5: warps 4-7 write the operand stages like the gather groups (one 128 x 64 stage per ~k block), bit 6: warps 8-11
  // read 64 accumulator columns every other k block like the accumulate warps — concurrent tensor-memory traffic
  if ((PAT & 32) && warp >= 4 && warp < 8) {
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 128u;
    uint32_t v[16];
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = threadIdx.x * 7919u + j;
    for (int it = 0; it < iters; it++) {
      const uint32_t ta = t_lane + 64u * (uint32_t)(it % 6);
      ptx::tmem_st_32x16(ta, v); ptx::tmem_st_32x16(ta + 16u, v); ptx::tmem_st_32x16(ta + 32u, v); ptx::tmem_st_32x16(ta + 48u, v);
      ptx::tmem_st_wait();
      __nanosleep(150);
    }
  }
  if ((PAT & 64) && warp >= 8 && warp < 12) {
    const uint32_t t0 = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    for (int it = 0; it < iters / 2; it++) {
      uint32_t r0[32];
      ptx::tmem_ld_32x32(t0 + (uint32_t)((it & 1) * 64), r0); ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i++) acc += __uint_as_float(r0[i]);
      ptx::tmem_ld_32x32(t0 + (uint32_t)((it & 1) * 64) + 32u, r0); ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i++) acc += __uint_as_float(r0[i]);
      __nanosleep(400);
    }
    if (acc == 1.2345f) asm volatile("st.shared.f32 [%0], %1;" ::"r"(base), "f"(acc));
  }
  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<1>(tmem, 512);
}

template <class F>
static int time_launch(F&& launch, int reps, float* best_ms) {
  cudaEvent_t e0, e1;
  AM_CUDA_TRY(cudaEventCreate(&e0));
  AM_CUDA_TRY(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < reps + 1; r++) {
    AM_CUDA_TRY(cudaEventRecord(e0, 0));
    launch();
    AM_CUDA_TRY(cudaEventRecord(e1, 0));
    AM_CUDA_TRY(cudaEventSynchronize(e1));
    AM_CUDA_TRY(cudaGetLastError());
    float ms = 0;
    AM_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 && ms < best) best = ms;   // first rep is warm-up
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *best_ms = best;
  return AM_OK;
}

int microbench(int which, double* tops) {
  if (!tops) { set_last_error("microbench: null output"); return AM_ERR_INVALID; }
  const int sms = sm_count();
  void* scratch = nullptr;
  int rc = workspace(kWsConv, (size_t)sms * 8 * 256 * 8, &scratch);
  if (rc) return rc;
  float ms = 0;
  const int blocks = sms * 8, threads = 256;
  double ops = 0;
  constexpr int CH = 12;
  switch (which) {
    case 0: {
      const int iters = 1 << 15;
      rc = time_launch([&] { simt_peak_kernel<float, CH><<<blocks, threads>>>((float*)scratch, iters, 1.0001f, 0.9999f); g_launch_count++; }, 3, &ms);
      ops = 2.0 * CH * (double)iters * blocks * threads;
    } break;
    case 1: {
      const int iters = 1 << 13;
      rc = time_launch([&] { simt_peak_kernel<double, CH><<<blocks, threads>>>((double*)scratch, iters, 1.0001, 0.9999); g_launch_count++; }, 3, &ms);
      ops = 2.0 * CH * (double)iters * blocks * threads;
    } break;
    case 2: {
      const int iters = 1 << 15;
      rc = time_launch([&] { simt_peak_kernel<int32_t, CH><<<blocks, threads>>>((int32_t*)scratch, iters, 3, 5); g_launch_count++; }, 3, &ms);
      ops = 2.0 * CH * (double)iters * blocks * threads;
    } break;
    case 3: {
      const int iters = 1 << 13;
      rc = time_launch([&] { simt_peak_kernel<int64_t, CH><<<blocks, threads>>>((int64_t*)scratch, iters, 0x100000003ll, 0x500000007ll); g_launch_count++; }, 3, &ms);
      ops = 2.0 * CH * (double)iters * blocks * threads;
    } break;
    case 4: {
      const int iters = 1 << 12;
      rc = time_launch([&] { dmma_peak_kernel<8><<<blocks, threads>>>((double*)scratch, iters); g_launch_count++; }, 3, &ms);
      ops = 2.0 * 256.0 * 8 * (double)iters * blocks * (threads / 32);
    } break;
    case 7: {
      const int iters = 1 << 14;
      rc = time_launch([&] { i64_narrow_peak_kernel<CH><<<blocks, threads>>>((int64_t*)scratch, iters, 3, 5); g_launch_count++; }, 3, &ms);
      ops = 2.0 * CH * (double)iters * blocks * threads;
    } break;
    case 13: {   // packed FFMA2: two FMAs per lane and instruction
      const int iters = 1 << 15;
      rc = time_launch([&] { ffma2_peak_kernel<CH><<<blocks, threads>>>((float*)scratch, iters, 1.0001f, 0.9999f); g_launch_count++; }, 3, &ms);
      ops = 4.0 * CH * (double)iters * blocks * threads;
    } break;
    case 20: case 21: case 23: case 27: case 31: case 32: case 33: case 34: case 35: {   // 33: + operand-stage writers, 34: + accumulator readers, 35: all   // conv issue pattern, see umma_pattern_kernel (32: all pieces + the weight stream)
      if (!gemm_f32_tc_available()) { set_last_error("microbench: tcgen05 needs compute capability 10.x"); return AM_ERR_UNSUPPORTED; }
      const int iters = 4096 * 4, smem = 32768 + 4 * 16384 + 1024;
      void* gbuf = nullptr;
      if ((rc = workspace(kWsMisc, 256 * 1024 + 1024, &gbuf))) return rc;
      const float* gsrc = (const float*)(((uintptr_t)gbuf + 127) & ~(uintptr_t)127);
      const void* kern = which == 20 ? (const void*)umma_pattern_kernel<0> : which == 21 ? (const void*)umma_pattern_kernel<1>
                       : which == 23 ? (const void*)umma_pattern_kernel<3> : which == 27 ? (const void*)umma_pattern_kernel<7>
                       : which == 31 ? (const void*)umma_pattern_kernel<15> : which == 32 ? (const void*)umma_pattern_kernel<31>
                       : which == 33 ? (const void*)umma_pattern_kernel<15 + 32> : which == 34 ? (const void*)umma_pattern_kernel<15 + 64>
                                                                                  : (const void*)umma_pattern_kernel<127>;
      AM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      rc = time_launch([&] {
        int it_arg = iters;
        void* args[2] = {&it_arg, (void*)&gsrc};
        cudaLaunchKernel(kern, dim3(sms), dim3(384), args, smem, 0);
        g_launch_count++;
      }, 3, &ms);
      ops = 2.0 * 128.0 * 64.0 * 8.0 * 12.0 * (double)iters * sms;
    } break;
    case 5:
    case 6:
    case 8: case 9: case 10: case 11: case 12: {   // 8..12: small-N variants (what the implicit-GEMM conv issues)
      if (!gemm_f32_tc_available()) { set_last_error("microbench: tcgen05 needs compute capability 10.x"); return AM_ERR_UNSUPPORTED; }
      const int cg = which == 6 ? 2 : 1;
      const int iters = 2000, batch = 16;
      // 8: N=64 one accumulator | 9: N=64 two accumulators | 10: N=64, A from TMEM | 11: N=64, A from TMEM, 2 acc | 12: N=128
      const int n_mma = (which >= 8 && which <= 11) ? 64 : (which == 12 ? 128 : 256);
      const int smem = 16384 + 32768 + 64 + 1024;
      const void* kern = nullptr;
      switch (which) {
        case 5: kern = (const void*)umma_peak_kernel<1, 256, 1, 0>; break;
        case 6: kern = (const void*)umma_peak_kernel<2, 256, 1, 0>; break;
        case 8: kern = (const void*)umma_peak_kernel<1, 64, 1, 0>; break;
        case 9: kern = (const void*)umma_peak_kernel<1, 64, 2, 0>; break;
        case 10: kern = (const void*)umma_peak_kernel<1, 64, 1, 1>; break;
        case 11: kern = (const void*)umma_peak_kernel<1, 64, 2, 1>; break;
        default: kern = (const void*)umma_peak_kernel<1, 128, 1, 0>; break;
      }
      AM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      const int grid = (sms / 2) * 2;
      rc = time_launch([&] {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cg; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int it_arg = iters;
        void* args[1] = {&it_arg};
        cudaLaunchKernelExC(&cfg, kern, args);
        g_launch_count++;
      }, 3, &ms);
      ops = 2.0 * (128.0 * cg) * (double)n_mma * 8.0 * (double)iters * batch * (grid / cg);
    } break;
    default:
      set_last_error("microbench: unknown selector %d", which);
      return AM_ERR_INVALID;
  }
  if (rc) return rc;
  *tops = ops / (ms * 1e-3) / 1e12;
  return AM_OK;
}

}  // namespace am
