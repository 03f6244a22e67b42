// Multi-GPU entry points of the C ABI (include/am_b200.h, "multi-GPU" section): ONE host process drives the GPUs of a
// box — what a Nim caller can use (the reference has no multi-device code and no cudaSetDevice anywhere, SURVEY F1 /
// §8b last row: "multi-GPU entry points manage devices themselves and restore the caller's device").  The
// one-process-per-GPU flavour (torch.distributed + symmetric memory) lives in arraymancer_b200/distributed.py and uses
// the same kernels.
//
// SURVEY §8e partitioning: rank g owns a contiguous block of rows of A and of C, B is replicated, no K split (integer
// results stay bit-exact, float results are those of the single-GPU kernels).
//   float32            the tcgen05 GEMM of every GPU stores its C tiles into EVERY GPU's copy of C from its epilogue
//                      (am_gemm_packed_f32_bcast over peer-mapped pointers: cudaDeviceEnablePeerAccess + UVA) — GEMM
//                      and all-gather are one kernel, NVLink traffic is spread over the mainloop;
//   float64 / integers every GPU computes its block in row chunks; a finished chunk is pushed to the peers by the copy
//                      engines (cudaMemcpy2DAsync on a second stream) while the next chunk computes.
// All entries only enqueue work (one stream per GPU, owned by the context) and end with a cross-GPU barrier made of
// events, so that am_mg_synchronize() — or any later call on the context — sees every copy of C complete.
#include <vector>

#include "am_common.cuh"
#include "gemm_dispatch.h"

struct am_mg_ctx {
  int ndev = 0;
  std::vector<int> dev;
  std::vector<cudaStream_t> st, st_copy;
  std::vector<cudaEvent_t> ev, ev_chunk;
  std::vector<void*> hA, hB;                 // cached packed-operand handles of the float32 path (per GPU)
  std::vector<int64_t> hA_rows, hA_k, hB_n, hB_k;
};

namespace am {

struct DeviceGuard {          // restores the caller's current device
  int prev = 0;
  DeviceGuard() { cudaGetDevice(&prev); }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

static void rows_of(int64_t M, int G, int g, int64_t* r0, int64_t* rows) {
  // contiguous blocks, multiples of 256 rows where possible (tile height of the tensor-core kernel), remainder last
  int64_t per = ((M + G - 1) / G + 255) / 256 * 256;
  if (per * (G - 1) >= M) per = (M + G - 1) / G;
  *r0 = per * g < M ? per * g : M;
  const int64_t end = (g == G - 1) ? M : (per * (g + 1) < M ? per * (g + 1) : M);
  *rows = end - *r0;
}

static int mg_barrier(am_mg_ctx* c) {
  // every stream records; every stream waits for all records
  for (int g = 0; g < c->ndev; g++) {
    AM_CUDA_TRY(cudaSetDevice(c->dev[g]));
    AM_CUDA_TRY(cudaEventRecord(c->ev[g], c->st[g]));
  }
  for (int g = 0; g < c->ndev; g++) {
    AM_CUDA_TRY(cudaSetDevice(c->dev[g]));
    for (int h = 0; h < c->ndev; h++)
      if (h != g) AM_CUDA_TRY(cudaStreamWaitEvent(c->st[g], c->ev[h], 0));
  }
  return AM_OK;
}

template <class T, class F>
static int mg_gemm_generic(am_mg_ctx* c, F&& gemm, int64_t M, int64_t N, int64_t K, T alpha, const T* const* A_local,
                           int64_t ldA, const T* const* B, int64_t ldB, T* const* C, int64_t ldC) {
  const int G = c->ndev;
  const int NCH = 4;
  int rc;
  if ((rc = mg_barrier(c))) return rc;                   // peers have finished with the previous contents of C
  for (int g = 0; g < G; g++) {
    int64_t r0, rows;
    rows_of(M, G, g, &r0, &rows);
    if (rows == 0) continue;
    AM_CUDA_TRY(cudaSetDevice(c->dev[g]));
    const int64_t chunk = (rows + NCH - 1) / NCH;
    for (int j = 0; j < NCH; j++) {
      const int64_t c0 = j * chunk, cr = (rows - c0 < chunk) ? rows - c0 : chunk;
      if (cr <= 0) break;
      rc = gemm(c->st[g], cr, N, K, alpha, A_local[g] + c0 * ldA, ldA, (int64_t)1, B[g], ldB, (int64_t)1, T(0),
                C[g] + (r0 + c0) * ldC, ldC, (int64_t)1);
      if (rc) return rc;
      if (G == 1) continue;
      cudaEvent_t e = c->ev_chunk[g * NCH + j];
      AM_CUDA_TRY(cudaEventRecord(e, c->st[g]));
      AM_CUDA_TRY(cudaStreamWaitEvent(c->st_copy[g], e, 0));
      for (int q = 1; q < G; q++) {                      // push the finished chunk to every peer (copy engines, NVLink)
        const int h = (g + q) % G;
        AM_CUDA_TRY(cudaMemcpy2DAsync(C[h] + (r0 + c0) * ldC, (size_t)ldC * sizeof(T), C[g] + (r0 + c0) * ldC, (size_t)ldC * sizeof(T),
                                      (size_t)N * sizeof(T), (size_t)cr, cudaMemcpyDeviceToDevice, c->st_copy[g]));
      }
    }
    if (G > 1) {                                          // the compute stream joins its copy stream
      AM_CUDA_TRY(cudaEventRecord(c->ev_chunk[g * NCH], c->st_copy[g]));
      AM_CUDA_TRY(cudaStreamWaitEvent(c->st[g], c->ev_chunk[g * NCH], 0));
    }
  }
  return mg_barrier(c);
}

}  // namespace am

using namespace am;

extern "C" {

int am_mg_init(int ndev, const int* devices, am_mg_ctx** out) {
  if (!out || ndev < 1 || ndev > 8) { set_last_error("am_mg_init: 1..8 devices"); return AM_ERR_INVALID; }
  int count = 0;
  AM_CUDA_TRY(cudaGetDeviceCount(&count));
  DeviceGuard guard;
  am_mg_ctx* c = new am_mg_ctx();
  c->ndev = ndev;
  for (int g = 0; g < ndev; g++) {
    const int d = devices ? devices[g] : g;
    if (d < 0 || d >= count) { delete c; set_last_error("am_mg_init: device %d does not exist (%d visible)", d, count); return AM_ERR_INVALID; }
    c->dev.push_back(d);
  }
  // peer access between every pair of distinct devices (already-enabled is fine)
  for (int g = 0; g < ndev; g++) {
    if (cudaSetDevice(c->dev[g]) != cudaSuccess) { delete c; return cuda_fail(cudaGetLastError(), "cudaSetDevice"); }
    for (int h = 0; h < ndev; h++) {
      if (c->dev[h] == c->dev[g]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, c->dev[g], c->dev[h]);
      if (!can) { delete c; set_last_error("am_mg_init: device %d cannot access device %d (no peer path)", c->dev[g], c->dev[h]); return AM_ERR_UNSUPPORTED; }
      cudaError_t e = cudaDeviceEnablePeerAccess(c->dev[h], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { delete c; return cuda_fail(e, "cudaDeviceEnablePeerAccess"); }
      cudaGetLastError();
    }
  }
  c->st.resize(ndev); c->st_copy.resize(ndev); c->ev.resize(ndev); c->ev_chunk.resize(ndev * 4);
  c->hA.assign(ndev, nullptr); c->hB.assign(ndev, nullptr);
  c->hA_rows.assign(ndev, 0); c->hA_k.assign(ndev, 0); c->hB_n.assign(ndev, 0); c->hB_k.assign(ndev, 0);
  for (int g = 0; g < ndev; g++) {
    cudaSetDevice(c->dev[g]);
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&c->st[g], cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->st_copy[g], cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->ev[g], cudaEventDisableTiming)) != cudaSuccess) { return cuda_fail(e, "am_mg_init: stream/event"); }
    for (int j = 0; j < 4; j++)
      if ((e = cudaEventCreateWithFlags(&c->ev_chunk[g * 4 + j], cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(e, "am_mg_init: event");
  }
  *out = c;
  return AM_OK;
}

int am_mg_device_count(const am_mg_ctx* c) { return c ? c->ndev : 0; }

void* am_mg_stream(const am_mg_ctx* c, int g) { return (c && g >= 0 && g < c->ndev) ? (void*)c->st[g] : nullptr; }

int am_mg_rows(const am_mg_ctx* c, int64_t M, int g, int64_t* row0, int64_t* rows) {
  if (!c || g < 0 || g >= c->ndev || M < 0 || !row0 || !rows) { set_last_error("am_mg_rows: bad argument"); return AM_ERR_INVALID; }
  rows_of(M, c->ndev, g, row0, rows);
  return AM_OK;
}

int am_mg_synchronize(am_mg_ctx* c) {
  if (!c) { set_last_error("am_mg_synchronize: null context"); return AM_ERR_INVALID; }
  DeviceGuard guard;
  for (int g = 0; g < c->ndev; g++) {
    AM_CUDA_TRY(cudaSetDevice(c->dev[g]));
    AM_CUDA_TRY(cudaStreamSynchronize(c->st_copy[g]));
    AM_CUDA_TRY(cudaStreamSynchronize(c->st[g]));
  }
  return AM_OK;
}

int am_mg_destroy(am_mg_ctx* c) {
  if (!c) return AM_OK;
  DeviceGuard guard;
  for (int g = 0; g < c->ndev; g++) {
    cudaSetDevice(c->dev[g]);
    cudaStreamSynchronize(c->st[g]); cudaStreamSynchronize(c->st_copy[g]);
    if (c->hA[g]) packed_free_f32(c->hA[g]);
    if (c->hB[g]) packed_free_f32(c->hB[g]);
    cudaStreamDestroy(c->st[g]); cudaStreamDestroy(c->st_copy[g]);
    cudaEventDestroy(c->ev[g]);
    for (int j = 0; j < 4; j++) cudaEventDestroy(c->ev_chunk[g * 4 + j]);
  }
  delete c;
  return AM_OK;
}

int am_mg_gemm_rowsharded_f32(am_mg_ctx* c, int64_t M, int64_t N, int64_t K, float alpha, const float* const* A_local,
                              int64_t ldA, const float* const* B, int64_t ldB, float* const* C, int64_t ldC) {
  if (!c || !A_local || !B || !C || M < 0 || N < 0 || K < 0 || ldA < K || ldB < N || ldC < N) {
    set_last_error("am_mg_gemm_rowsharded_f32: bad argument"); return AM_ERR_INVALID;
  }
  if (M == 0 || N == 0 || K == 0) return AM_OK;
  DeviceGuard guard;
  const int G = c->ndev;
  // small / skinny blocks: the generic path (am_gemm_strided_f32 picks the kernel) + copy-engine pushes
  int64_t r0, rows;
  rows_of(M, G, 0, &r0, &rows);
  const bool tc = gemm_f32_tc_available() && rows >= 256 && N >= 256 && K >= 256;
  if (!tc) {
    return mg_gemm_generic<float>(c, [](cudaStream_t s, int64_t m, int64_t n, int64_t k, float al, const float* a, int64_t ra, int64_t ca,
                                        const float* b, int64_t rb, int64_t cb, float be, float* cc, int64_t rc_, int64_t cs) {
      return am_gemm_strided_f32((am_stream_t)s, m, n, k, al, a, ra, ca, b, rb, cb, be, cc, rc_, cs);
    }, M, N, K, alpha, A_local, ldA, B, ldB, C, ldC);
  }
  int rc;
  if ((rc = mg_barrier(c))) return rc;                    // peers have finished with the previous contents of C
  for (int g = 0; g < G; g++) {
    rows_of(M, G, g, &r0, &rows);
    if (rows == 0) continue;
    AM_CUDA_TRY(cudaSetDevice(c->dev[g]));
    // packed operands are cached in the context and refreshed in place while the shapes repeat
    if (c->hA[g] && (c->hA_rows[g] != rows || c->hA_k[g] != K)) { packed_free_f32(c->hA[g]); c->hA[g] = nullptr; }
    if (c->hB[g] && (c->hB_n[g] != N || c->hB_k[g] != K)) { packed_free_f32(c->hB[g]); c->hB[g] = nullptr; }
    if (!c->hA[g]) rc = pack_f32(c->st[g], rows, K, A_local[g], ldA, 1, &c->hA[g]);
    else rc = repack_f32(c->st[g], c->hA[g], A_local[g], ldA, 1);
    if (rc) return rc;
    c->hA_rows[g] = rows; c->hA_k[g] = K;
    if (!c->hB[g]) rc = pack_f32(c->st[g], N, K, B[g], 1, ldB, &c->hB[g]);
    else rc = repack_f32(c->st[g], c->hB[g], B[g], 1, ldB);
    if (rc) return rc;
    c->hB_n[g] = N; c->hB_k[g] = K;
    float* peers[8];
    for (int h = 0; h < G; h++) peers[h] = C[h] + r0 * ldC;
    if ((rc = gemm_packed_f32_bcast(c->st[g], alpha, c->hA[g], c->hB[g], G, peers, g, ldC, 1))) return rc;
  }
  return mg_barrier(c);
}

#define DEF_MG_GENERIC(SUF, T)                                                                                              \
  int am_mg_gemm_rowsharded_##SUF(am_mg_ctx* c, int64_t M, int64_t N, int64_t K, T alpha, const T* const* A_local, int64_t ldA, \
                                  const T* const* B, int64_t ldB, T* const* C, int64_t ldC) {                                \
    if (!c || !A_local || !B || !C || M < 0 || N < 0 || K < 0 || ldA < K || ldB < N || ldC < N) {                          \
      set_last_error("am_mg_gemm_rowsharded: bad argument"); return AM_ERR_INVALID;                                        \
    }                                                                                                                       \
    if (M == 0 || N == 0 || K == 0) return AM_OK;                                                                           \
    DeviceGuard guard;                                                                                                      \
    return mg_gemm_generic<T>(c, [](cudaStream_t s, int64_t m, int64_t n, int64_t k, T al, const T* a, int64_t ra, int64_t ca, \
                                    const T* b, int64_t rb, int64_t cb, T be, T* cc, int64_t rc_, int64_t cs) {             \
      return am_gemm_strided_##SUF((am_stream_t)s, m, n, k, al, a, ra, ca, b, rb, cb, be, cc, rc_, cs);                     \
    }, M, N, K, alpha, A_local, ldA, B, ldB, C, ldC);                                                                       \
  }
DEF_MG_GENERIC(f64, double)
DEF_MG_GENERIC(i32, int32_t)
DEF_MG_GENERIC(i64, int64_t)

// Host-buffer product over all GPUs of the context: A[M,K], B[K,N], C[M,N] row-major in HOST memory (pinned for
// full-speed copies).  GPU g uploads its rows of A and only its 1/G share of B's rows; the shares are exchanged between
// the GPUs by the copy engines over NVLink; every GPU multiplies its block (tcgen05 for large shapes) and sends its rows
// of C back.  Synchronous: C is complete when the call returns.
int am_mg_host_gemm_f32(am_mg_ctx* c, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t ldA, const float* B,
                        int64_t ldB, float* C, int64_t ldC) {
  if (!c || !A || !B || !C || M < 0 || N < 0 || K < 0 || ldA < K || ldB < N || ldC < N) {
    set_last_error("am_mg_host_gemm_f32: bad argument"); return AM_ERR_INVALID;
  }
  if (M == 0 || N == 0 || K == 0) return AM_OK;
  DeviceGuard guard;
  const int G = c->ndev;
  std::vector<float*> dA(G, nullptr), dB(G, nullptr), dC(G, nullptr);
  int status = AM_OK;
  cudaError_t e = cudaSuccess;
  auto fail = [&](cudaError_t err, const char* what) { if (!status) status = cuda_fail(err, what); };
  // B's rows are dealt in G contiguous shares
  auto kshare = [&](int g, int64_t* k0, int64_t* kn) { const int64_t per = (K + G - 1) / G; *k0 = per * g < K ? per * g : K; *kn = (K - *k0 < per) ? K - *k0 : per; };
  for (int g = 0; g < G && !status; g++) {
    int64_t r0, rows, k0, kn;
    rows_of(M, G, g, &r0, &rows); kshare(g, &k0, &kn);
    cudaSetDevice(c->dev[g]);
    // plain cudaMalloc: peers read dB through their own copy engines (stream-ordered pool memory is not peer-visible by default)
    if ((e = cudaMalloc((void**)&dB[g], (size_t)(K * N) * 4)) != cudaSuccess) { fail(e, "cudaMalloc B"); break; }
    if (rows > 0) {
      if ((e = cudaMalloc((void**)&dA[g], (size_t)(rows * K) * 4)) != cudaSuccess ||
          (e = cudaMalloc((void**)&dC[g], (size_t)(rows * N) * 4)) != cudaSuccess) { fail(e, "cudaMalloc A/C"); break; }
    }
    if (kn > 0 && (e = cudaMemcpy2DAsync(dB[g] + k0 * N, (size_t)N * 4, B + k0 * ldB, (size_t)ldB * 4, (size_t)N * 4, (size_t)kn,
                                         cudaMemcpyHostToDevice, c->st[g])) != cudaSuccess) { fail(e, "H2D B share"); break; }
    if (rows > 0 && (e = cudaMemcpy2DAsync(dA[g], (size_t)K * 4, A + r0 * ldA, (size_t)ldA * 4, (size_t)K * 4, (size_t)rows,
                                           cudaMemcpyHostToDevice, c->st[g])) != cudaSuccess) { fail(e, "H2D A rows"); break; }
  }
  if (!status) status = mg_barrier(c);                    // every share of B is on its GPU (and every dB is allocated)
  for (int g = 0; g < G && !status; g++) {
    cudaSetDevice(c->dev[g]);
    for (int q = 1; q < G && !status; q++) {              // pull the other shares (copy engines over NVLink)
      const int h = (g + q) % G;
      int64_t k0, kn;
      kshare(h, &k0, &kn);
      if (kn > 0 && (e = cudaMemcpyAsync(dB[g] + k0 * N, dB[h] + k0 * N, (size_t)(kn * N) * 4, cudaMemcpyDeviceToDevice, c->st[g])) != cudaSuccess)
        fail(e, "peer copy of a B share");
    }
    int64_t r0, rows;
    rows_of(M, G, g, &r0, &rows);
    if (rows > 0 && !status) {
      status = am_gemm_strided_f32((am_stream_t)c->st[g], rows, N, K, alpha, dA[g], K, 1, dB[g], N, 1, 0.f, dC[g], N, 1);
      if (!status && (e = cudaMemcpy2DAsync(C + r0 * ldC, (size_t)ldC * 4, dC[g], (size_t)N * 4, (size_t)N * 4, (size_t)rows,
                                            cudaMemcpyDeviceToHost, c->st[g])) != cudaSuccess) fail(e, "D2H C rows");
    }
  }
  // peers may still be reading this GPU's share of B: barrier before the buffers are released
  if (mg_barrier(c) != AM_OK && !status) status = AM_ERR_CUDA;
  for (int g = 0; g < G; g++) {
    cudaSetDevice(c->dev[g]);
    if ((e = cudaStreamSynchronize(c->st[g])) != cudaSuccess) fail(e, "sync");
    if (dA[g]) cudaFree(dA[g]);
    if (dB[g]) cudaFree(dB[g]);
    if (dC[g]) cudaFree(dC[g]);
  }
  return status;
}

}  // extern "C"
