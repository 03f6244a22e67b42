// float64 GEMM on the FP64 tensor pipe: mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4).
//
// tcgen05.mma has no FP64 kind (kinds: tf32/f16/i8/f8f6f4/mx*), so the B200 FP64 tensor path is
// still the warp-level DMMA.  Measured on B200 (profiles/r01_bringup.md): DMMA peak
// 37.0 TFLOP/s vs DFMA 34.8 TFLOP/s — and the DFMA register-tile kernel only reached 52 % of its
// peak (three 64-bit register operands per FMA), while one DMMA does 256 FMAs per issue slot.
//
// Same structure as the SIMT core (contract_simt.cuh): operands come through the strided loaders
// (any row/col stride), are staged via registers into padded shared memory (double-buffered), and
// each of the 8 warps owns a 64x32 slice of the 128x128 CTA tile = 8x4 DMMA tiles, accumulators
// (64 doubles/lane) in registers.  Accumulation order per output element: k ascending in groups of
// 4 — within the stated 1e-13 relative tolerance of the reference's sequential order.
#include <cuda.h>

#include "contract_simt.cuh"
#include "gemm_dispatch.h"
#include "ptx_sm100.cuh"

namespace am {

int make_tmap_f64_2d(CUtensorMap* tm, const double* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                     uint32_t box_inner, uint32_t box_outer);     // gemm_f32_tc.cu

struct DmmaCfg {
  static constexpr int BM = 128, BN = 128, BK = 8, NT = 256;
  static constexpr int LDA = BM + 4, LDB = BN + 4;    // +4 doubles: (4k + m) mod 16 distinct -> conflict-free fragment loads
  static constexpr int EA = BM * BK / NT, EB = BN * BK / NT;
  static constexpr int WM = 64, WN = 32;              // warp tile
};

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <class LA, class LB, class Epi>
__global__ void __launch_bounds__(DmmaCfg::NT, 1)
contract_dmma_kernel(const LA la, const LB lb, const Epi epi, int64_t K, int a_kfast, int b_kfast) {
  using Cfg = DmmaCfg;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, NT = Cfg::NT, LDA = Cfg::LDA, LDB = Cfg::LDB;
  constexpr int EA = Cfg::EA, EB = Cfg::EB;
  __shared__ __align__(16) double As[2][BK][LDA];
  __shared__ __align__(16) double Bs[2][BK][LDB];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp >> 2) * Cfg::WM, wn0 = (warp & 3) * Cfg::WN;
  const int fk = lane & 3, fr = lane >> 2;           // fragment coordinates: k within the group of 4, row/col within 8
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;

  const bool akf = LA::kMapping == 0 ? (a_kfast != 0) : (LA::kMapping == 2);
  const bool bkf = LB::kMapping == 0 ? (b_kfast != 0) : (LB::kMapping == 2);
  typename LA::Slot sa[EA];
  typename LB::Slot sb[EB];
  int oa[EA], ob[EB];
#pragma unroll
  for (int i = 0; i < EA; i++) {
    const int idx = tid + i * NT;
    const int mi = akf ? idx / BK : idx % BM, ki = akf ? idx % BK : idx / BM;
    sa[i] = la.slot(m0 + mi, ki);
    oa[i] = ki * LDA + mi;
  }
#pragma unroll
  for (int i = 0; i < EB; i++) {
    const int idx = tid + i * NT;
    const int ni = bkf ? idx / BK : idx % BN, ki = bkf ? idx % BK : idx / BN;
    sb[i] = lb.slot(n0 + ni, ki);
    ob[i] = ki * LDB + ni;
  }

  double c[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) c[i][j][0] = c[i][j][1] = 0.0;

  double ra[EA], rb[EB];
  const int64_t ntiles = (K + BK - 1) / BK;
  if (ntiles > 0) {
#pragma unroll
    for (int i = 0; i < EA; i++) ra[i] = la.load(sa[i], 0);
#pragma unroll
    for (int i = 0; i < EB; i++) rb[i] = lb.load(sb[i], 0);
#pragma unroll
    for (int i = 0; i < EA; i++) (&As[0][0][0])[oa[i]] = ra[i];
#pragma unroll
    for (int i = 0; i < EB; i++) (&Bs[0][0][0])[ob[i]] = rb[i];
  }
  __syncthreads();

  for (int64_t t = 0; t < ntiles; t++) {
    const int buf = (int)(t & 1);
    const bool more = (t + 1 < ntiles);
    if (more) {
      const int64_t kt = (t + 1) * BK;
#pragma unroll
      for (int i = 0; i < EA; i++) ra[i] = la.load(sa[i], kt);
#pragma unroll
      for (int i = 0; i < EB; i++) rb[i] = lb.load(sb[i], kt);
    }
#pragma unroll
    for (int k4 = 0; k4 < BK; k4 += 4) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = As[buf][k4 + fk][wm0 + i * 8 + fr];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[buf][k4 + fk][wn0 + j * 8 + fr];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma_m8n8k4(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
    if (more) {
#pragma unroll
      for (int i = 0; i < EA; i++) (&As[buf ^ 1][0][0])[oa[i]] = ra[i];
#pragma unroll
      for (int i = 0; i < EB; i++) (&Bs[buf ^ 1][0][0])[ob[i]] = rb[i];
    }
    __syncthreads();
  }

  // C fragment: lane holds C[row = fr][cols = 2*fk, 2*fk+1] of every 8x8 tile
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int64_t m = m0 + wm0 + i * 8 + fr;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const double out[2] = {c[i][j][0], c[i][j][1]};
      epi.template store<2>(m, n0 + wn0 + j * 8 + 2 * fk, out, 0);
    }
  }
}

// ------------------------------------------------------------------ TMA-fed variant (unit-stride operands)
// Operand tiles come in by cp.async.bulk.tensor (2-D boxes, zero fill outside the matrix = free edge handling) through a
// 5-stage full / empty mbarrier ring: no staging instructions, no __syncthreads in the mainloop.  A box is laid out the way
// the operand is stored: k-contiguous operands land as [128 rows][BK], mn-contiguous ones as [BK][128]; the DMMA fragment
// loads index either (template flags).  The dense boxes cost 2-way bank conflicts on the 64-bit fragment loads, which is
// irrelevant next to a DMMA issue interval of 16 cycles per scheduler.  Thread 0 is the producer (prologue of STAGES-1
// tiles, then one refill per loop trip); 16 warps x (32x32) warp tiles consume: four per scheduler keep the FP64 tensor
// pipe busy while others wait for a stage (the 8-warp register-staged kernel above ran
// two per scheduler and two barriers per 8-deep tile: 0.79 of the DMMA peak).
struct DmmaTmaCfg {
  static constexpr int BM = 128, BN = 128, BK = 16, NT = 512, STAGES = 5;
  static constexpr int TILE_BYTES = 128 * BK * 8;                  // one operand tile
  static constexpr int STAGE_BYTES = 2 * TILE_BYTES;               // 32 KB
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 128 + 64;
};

template <bool A_KMAJ, bool B_KMAJ, class Epi>
__global__ void __launch_bounds__(DmmaTmaCfg::NT, 1)
contract_dmma_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Epi epi, int64_t K) {
  using Cfg = DmmaTmaCfg;
  constexpr int BK = Cfg::BK, S = Cfg::STAGES;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 127u) & ~127u;
  const double* sm = reinterpret_cast<const double*>(smem_raw + (base - ptx::smem_u32(smem_raw)));
  const uint32_t bar0 = base + S * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;
  const int fk = lane & 3, fr = lane >> 2;
  const int m0 = blockIdx.y * Cfg::BM, n0 = blockIdx.x * Cfg::BN;
  if (tid == 0) {
    ptx::prefetch_tensormap(&tmA); ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < S; s++) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), Cfg::NT / 32); }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int ntiles = (int)((K + BK - 1) / BK);
  auto issue = [&](int t) {                        // thread 0 only
    const int s = t % S;
    const uint32_t dst = base + (uint32_t)s * Cfg::STAGE_BYTES;
    ptx::mbar_arrive_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
    if (A_KMAJ) ptx::tma_load_2d(dst, &tmA, full_bar(s), t * BK, m0); else ptx::tma_load_2d(dst, &tmA, full_bar(s), m0, t * BK);
    if (B_KMAJ) ptx::tma_load_2d(dst + Cfg::TILE_BYTES, &tmB, full_bar(s), t * BK, n0);
    else ptx::tma_load_2d(dst + Cfg::TILE_BYTES, &tmB, full_bar(s), n0, t * BK);
  };
  if (tid == 0)
    for (int t = 0; t < S - 1 && t < ntiles; t++) issue(t);

  double c[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) c[i][j][0] = c[i][j][1] = 0.0;

  for (int t = 0; t < ntiles; t++) {
    const int s = t % S;
    ptx::mbar_wait(full_bar(s), (uint32_t)((t / S) & 1));
    // DELAYED release: the stage of tile t-1 is handed back only now, one loop trip after its last fragment load.  All
    // DMMAs of tile t-1 were issued before this point (in-order issue, the arrive cannot be hoisted across the loop
    // branch), and a DMMA issues only once its operands have arrived — so every shared-memory read of that stage has
    // completed.  (Releasing at the end of the same trip let ptxas place the arrive right behind the last LDS, ahead
    // of the DMMAs that consume it: the TMA refill could then overtake reads still in flight — observed as run-to-run
    // differences at K >= 4096.)
    if (t >= 1 && lane == 0) ptx::mbar_arrive(empty_bar((t - 1) % S));
    if (tid == 0 && t >= 1 && t + S - 2 < ntiles) {
      // tile t+S-2 goes into the stage tile t-2 occupied, released by every warp at the start of its trip t-1
      const int tn = t + S - 2;
      ptx::mbar_wait(empty_bar(tn % S), (uint32_t)((tn / S - 1) & 1));
      issue(tn);
    }
    const double* As = sm + (size_t)s * (Cfg::STAGE_BYTES / 8);
    const double* Bs = As + Cfg::TILE_BYTES / 8;
#pragma unroll
    for (int k4 = 0; k4 < BK; k4 += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int m = wm0 + i * 8 + fr;
        a[i] = A_KMAJ ? As[m * BK + k4 + fk] : As[(k4 + fk) * 128 + m];
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int n = wn0 + j * 8 + fr;
        b[j] = B_KMAJ ? Bs[n * BK + k4 + fk] : Bs[(k4 + fk) * 128 + n];
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma_m8n8k4(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int64_t m = (int64_t)m0 + wm0 + i * 8 + fr;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const double out[2] = {c[i][j][0], c[i][j][1]};
      epi.template store<2>(m, (int64_t)n0 + wn0 + j * 8 + 2 * fk, out, 0);
    }
  }
}

// operand X (mn x k view, strides in elements): can a 2-D tensor map describe it?  kmaj: unit stride along k.
static bool tma_operand_ok(const double* p, int64_t mn_stride, int64_t k_stride, int64_t MN, int64_t K, bool* kmaj) {
  if ((reinterpret_cast<uintptr_t>(p) & 15) != 0 || MN >= (1ll << 31) || K >= (1ll << 31)) return false;
  if (k_stride == 1 && mn_stride >= K && mn_stride % 2 == 0) { *kmaj = true; return true; }
  if (mn_stride == 1 && k_stride >= MN && k_stride % 2 == 0) { *kmaj = false; return true; }
  return false;
}

template <bool AK, bool BK_>
static int launch_dmma_tma(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB, const StridedEpilogue<double>& epi, int64_t M,
                           int64_t N, int64_t K) {
  auto kern = contract_dmma_tma_kernel<AK, BK_, StridedEpilogue<double>>;
  AM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DmmaTmaCfg::SMEM_BYTES));
  dim3 grid((unsigned)ceil_div(N, DmmaTmaCfg::BN), (unsigned)ceil_div(M, DmmaTmaCfg::BM), 1);
  kern<<<grid, DmmaTmaCfg::NT, DmmaTmaCfg::SMEM_BYTES, st>>>(tmA, tmB, epi, K);
  g_launch_count++;
  AM_CUDA_TRY(cudaGetLastError());
  return AM_OK;
}

int gemm_f64_dmma(cudaStream_t st, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t rsA,
                  int64_t csA, const double* B, int64_t rsB, int64_t csB, double beta, double* C, int64_t rsC,
                  int64_t csC, const double* bias_col) {
  int64_t a_mn = rsA, a_k = csA, b_mn = csB, b_k = rsB;
  const double *bias_m = nullptr, *bias_n = bias_col;
  if (iabs64(rsC) < iabs64(csC)) {
    bias_m = bias_col; bias_n = nullptr;    // column-major-ish C: compute C^T = B^T A^T (pairs of columns are the vector dim)
    const double* tp = A; A = B; B = tp;
    int64_t t;
    t = a_mn; a_mn = b_mn; b_mn = t;
    t = a_k; a_k = b_k; b_k = t;
    t = M; M = N; N = t;
    t = rsC; rsC = csC; csC = t;
  }
  using L = StridedLoader<double>;
  using Epi = StridedEpilogue<double>;
  const bool vec_ok = (csC == 1) && (rsC % 2 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  // TMA-fed kernel when both operands have a unit stride (and the grid fits): the common row-/column-major cases
  bool akm = false, bkm = false;
  if (tuning(kTuneDmmaTma) && tma_operand_ok(A, a_mn, a_k, M, K, &akm) && tma_operand_ok(B, b_mn, b_k, N, K, &bkm) &&
      ceil_div(M, DmmaTmaCfg::BM) <= 65535 && gemm_f32_tc_available()) {
    CUtensorMap tmA, tmB;
    int rc;
    if (akm) rc = make_tmap_f64_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)a_mn * 8, DmmaTmaCfg::BK, 128);
    else rc = make_tmap_f64_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)a_k * 8, 128, DmmaTmaCfg::BK);
    if (rc) return rc;
    if (bkm) rc = make_tmap_f64_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)b_mn * 8, DmmaTmaCfg::BK, 128);
    else rc = make_tmap_f64_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)b_k * 8, 128, DmmaTmaCfg::BK);
    if (rc) return rc;
    Epi epi{C, rsC, csC, M, N, alpha, beta, vec_ok, bias_m, bias_n};
    if (akm && bkm) return launch_dmma_tma<true, true>(st, tmA, tmB, epi, M, N, K);
    if (akm) return launch_dmma_tma<true, false>(st, tmA, tmB, epi, M, N, K);
    if (bkm) return launch_dmma_tma<false, true>(st, tmA, tmB, epi, M, N, K);
    return launch_dmma_tma<false, false>(st, tmA, tmB, epi, M, N, K);
  }
  const int64_t band = (int64_t)65535 * DmmaCfg::BM;
  for (int64_t r = 0; r < M; r += band) {
    const int64_t mb = (M - r < band) ? M - r : band;
    L la{A + r * a_mn, a_mn, a_k, mb, K};
    L lb{B, b_mn, b_k, N, K};
    Epi epi{C + r * rsC, rsC, csC, mb, N, alpha, beta, vec_ok, bias_m ? bias_m + r : nullptr, bias_n};
    dim3 grid((unsigned)ceil_div(N, DmmaCfg::BN), (unsigned)ceil_div(mb, DmmaCfg::BM), 1);
    contract_dmma_kernel<L, L, Epi><<<grid, DmmaCfg::NT, 0, st>>>(la, lb, epi, K, iabs64(a_k) <= iabs64(a_mn),
                                                                  iabs64(b_k) <= iabs64(b_mn));
    g_launch_count++;
    AM_CUDA_TRY(cudaGetLastError());
  }
  return AM_OK;
}

}  // namespace am
