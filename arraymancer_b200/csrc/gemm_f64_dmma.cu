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
#include "contract_simt.cuh"
#include "gemm_dispatch.h"

namespace am {

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

int gemm_f64_dmma(cudaStream_t st, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t rsA,
                  int64_t csA, const double* B, int64_t rsB, int64_t csB, double beta, double* C, int64_t rsC,
                  int64_t csC) {
  int64_t a_mn = rsA, a_k = csA, b_mn = csB, b_k = rsB;
  if (iabs64(rsC) < iabs64(csC)) {    // column-major-ish C: compute C^T = B^T A^T (pairs of columns are the vector dim)
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
  const int64_t band = (int64_t)65535 * DmmaCfg::BM;
  for (int64_t r = 0; r < M; r += band) {
    const int64_t mb = (M - r < band) ? M - r : band;
    L la{A + r * a_mn, a_mn, a_k, mb, K};
    L lb{B, b_mn, b_k, N, K};
    Epi epi{C + r * rsC, rsC, csC, mb, N, alpha, beta, vec_ok};
    dim3 grid((unsigned)ceil_div(N, DmmaCfg::BN), (unsigned)ceil_div(mb, DmmaCfg::BM), 1);
    contract_dmma_kernel<L, L, Epi><<<grid, DmmaCfg::NT, 0, st>>>(la, lb, epi, K, iabs64(a_k) <= iabs64(a_mn),
                                                                  iabs64(b_k) <= iabs64(b_mn));
    g_launch_count++;
    AM_CUDA_TRY(cudaGetLastError());
  }
  return AM_OK;
}

}  // namespace am
