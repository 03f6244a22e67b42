// extern "C" entry points of libarraymancer_b200.so (declared in include/am_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include <vector>

#include "am_common.cuh"
#include "gemm_dispatch.h"

namespace am {

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_last_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return AM_ERR_CUDA;
}

// ------------------------------------------------------------------ per-device workspace
struct WsEntry { void* p = nullptr; size_t bytes = 0; };
static std::mutex g_ws_mu;
static WsEntry g_ws[16][kWsNumSlots];
static int g_sm[16] = {0};
void conv_tables_invalidate();

int workspace(int slot, size_t bytes, void** out) {
  int dev = 0;
  AM_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16 || slot < 0 || slot >= kWsNumSlots) { set_last_error("workspace: bad slot"); return AM_ERR_INVALID; }
  std::lock_guard<std::mutex> lk(g_ws_mu);
  WsEntry& e = g_ws[dev][slot];
  if (e.bytes < bytes || !e.p) {
    if (e.p) {
      // the old block may still be in use by enqueued work: cudaFree synchronises the device first
      AM_CUDA_TRY(cudaFree(e.p));
      e.p = nullptr; e.bytes = 0;
      if (slot == kWsConvTab) conv_tables_invalidate();
    }
    size_t want = bytes < 256 ? 256 : bytes;
    want = (want + (want >> 3) + 255) & ~(size_t)255;          // 12.5% slack against regrowth
    AM_CUDA_TRY(cudaMalloc(&e.p, want));
    e.bytes = want;
  }
  *out = e.p;
  return AM_OK;
}
void workspace_release_all() {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < 16; d++)
    for (int s = 0; s < kWsNumSlots; s++)
      if (g_ws[d][s].p) {
        cudaSetDevice(d);
        cudaFree(g_ws[d][s].p);
        g_ws[d][s] = WsEntry{};
      }
  cudaSetDevice(cur);
  conv_tables_invalidate();
}
int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 148;
  if (!g_sm[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_sm[dev] = n;
  }
  return g_sm[dev];
}

static std::atomic<int> g_f32_path{AM_F32_AUTO};
static std::atomic<int> g_f64_path{AM_F64_AUTO};

// ------------------------------------------------------------------ explicit tuning knobs (no environment variables)
static std::atomic<int> g_tune[kTuneCount] = {{2}, {8}, {1}, {0}, {0}, {0}, {0}, {0}, {1}, {1}, {1}, {4}, {1}};
static const char* const kTuneNames[kTuneCount] = {"tc_flush_kb", "tc_group", "tc_sync", "pack_scalar", "host_rowchunks",
                                                   "convtc_groups", "convtc_debug", "convtc_dgrad_gather", "simt_vec_load", "dmma_tma", "convtc_hi_resident", "convtc_flush_kb", "convtc_wgrad_tma"};
int tuning(int key) { return (key >= 0 && key < kTuneCount) ? g_tune[key].load(std::memory_order_relaxed) : 0; }
static int tune_key(const char* name) {
  if (!name) return -1;
  for (int i = 0; i < kTuneCount; i++) if (strcmp(name, kTuneNames[i]) == 0) return i;
  return -1;
}

// ------------------------------------------------------------------ host-buffer entries: streams, events, memory pool
// The host-buffer GEMMs keep, per calling thread and device, three non-blocking streams (H2D, compute, D2H) and a
// growing list of timing-less events, and allocate their device buffers from ONE private stream-ordered pool per
// device (not the device's default pool, which the host framework may share): freed blocks stay cached in it between
// calls and are returned to the driver by am_shutdown().
struct HostCtx {
  int dev = -1;
  cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> events;
  size_t used = 0;
  ~HostCtx() { release(); }
  void release() {
    for (auto e : events) cudaEventDestroy(e);
    events.clear();
    if (s_in) cudaStreamDestroy(s_in);
    if (s_cmp) cudaStreamDestroy(s_cmp);
    if (s_out) cudaStreamDestroy(s_out);
    s_in = s_cmp = s_out = nullptr; dev = -1;
  }
  int init(int device) {
    used = 0;
    if (dev == device && s_in) return AM_OK;
    release();
    AM_CUDA_TRY(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    AM_CUDA_TRY(cudaStreamCreateWithFlags(&s_cmp, cudaStreamNonBlocking));
    AM_CUDA_TRY(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    dev = device;
    return AM_OK;
  }
  cudaEvent_t event() {           // next cached event (nullptr on failure)
    if (used == events.size()) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
      events.push_back(e);
    }
    return events[used++];
  }
};
static thread_local HostCtx t_host;

static std::mutex g_pool_mu;
static cudaMemPool_t g_pool[16] = {nullptr};
static int host_pool(int dev, cudaMemPool_t* out) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (dev < 0 || dev >= 16) { set_last_error("host gemm: device index out of range"); return AM_ERR_INVALID; }
  if (!g_pool[dev]) {
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    AM_CUDA_TRY(cudaMemPoolCreate(&g_pool[dev], &props));
    uint64_t thr = UINT64_MAX;                     // keep freed blocks cached between calls (trimmed by am_shutdown)
    AM_CUDA_TRY(cudaMemPoolSetAttribute(g_pool[dev], cudaMemPoolAttrReleaseThreshold, &thr));
  }
  *out = g_pool[dev];
  return AM_OK;
}
static void host_pools_release() {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  for (int d = 0; d < 16; d++)
    if (g_pool[d]) { cudaMemPoolDestroy(g_pool[d]); g_pool[d] = nullptr; }
}

template <class T>
static int check_gemm_args(int64_t M, int64_t N, int64_t K, const T* A, const T* B, T* C) {
  if (M < 0 || N < 0 || K < 0) { set_last_error("gemm_strided: negative dimension"); return AM_ERR_INVALID; }
  if (M == 0 || N == 0 || K == 0) return -1;   // nothing to do (K == 0 leaves C untouched, gemm.nim:203)
  if (!A || !B || !C) { set_last_error("gemm_strided: null operand pointer"); return AM_ERR_INVALID; }
  return AM_OK;
}

// f32: tcgen05 3xTF32 when the shape fills tensor tiles, the DRAM-streaming kernel for skinny products, exact FFMA kernel
// otherwise.  bias_col (nullable): the linear layer's bias, fused into the epilogue of whichever kernel runs.
int gemm_dispatch_f32(cudaStream_t st, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t rsA, int64_t csA,
                      const float* B, int64_t rsB, int64_t csB, float beta, float* C, int64_t rsC, int64_t csC, const float* bias_col) {
  const int path = g_f32_path.load();
  if (path == AM_F32_TC) return gemm_f32_tc(st, 2, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, nullptr, bias_col);
  if (path == AM_F32_TC_1CTA) return gemm_f32_tc(st, 1, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, nullptr, bias_col);
  if (path == AM_F32_AUTO && gemm_f32_tc_available()) {
    // worth it once K amortises the split/pack pre-pass (an extra pass over A and B) and there are a few output
    // tiles: measured crossover vs the SIMT kernel is around 3 GFLOP (LeNet's 4096x800x500 linear layer: 0.084 vs
    // 0.137 ms forward, 0.29 vs 0.385 ms backward; profiles/r01_bringup.md)
    const double flops = 2.0 * (double)M * (double)N * (double)K;
    if (M >= 256 && N >= 256 && K >= 256 && flops >= 3.0e9)
      return gemm_f32_tc(st, 2, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, nullptr, bias_col);
  }
  if (path == AM_F32_AUTO && !bias_col) {
    bool done = false;
    int rc = gemm_skinny<float>(st, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, &done);
    if (rc || done) return rc;
  }
  return gemm_simt<float>(st, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, bias_col);
}

int gemm_dispatch_f64(cudaStream_t st, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t rsA, int64_t csA,
                      const double* B, int64_t rsB, int64_t csB, double beta, double* C, int64_t rsC, int64_t csC, const double* bias_col) {
  const int path = g_f64_path.load();
  if (path == AM_F64_AUTO && !bias_col) {
    bool done = false;
    int rc = gemm_skinny<double>(st, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, &done);
    if (rc || done) return rc;
  }
  // DMMA kernel has one 128x128 tile shape: use it once those tiles cover most of the chip
  const bool big = ceil_div(M, 128) * ceil_div(N, 128) >= (2 * (int64_t)sm_count()) / 3;
  if (path == AM_F64_DMMA || (path == AM_F64_AUTO && big))
    return gemm_f64_dmma(st, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, bias_col);
  return gemm_simt<double>(st, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, bias_col);
}


// Large row-major float32 products from HOST buffers: K-pipelined schedule.  With row chunks alone (host_gemm below)
// the first product cannot start before all of B is on the device (a third of the whole PCIe time); splitting K makes
// the work that is computable grow LINEARLY with the bytes that have arrived:
//   phase 1  the first half of K in 4 slices: slice c of A (M x kc, 2-D copy) and of B (kc x N, contiguous rows) is
//            uploaded, split/packed and multiplied into the full C (beta = 0, then 1) while slice c+1 uploads;
//   phase 2  by the time phase 1 has computed, the second half of K has arrived: it is processed by ROW chunks
//            (C_i += A_i[:, K1:] * B[K1:, :]), each finished chunk of C going back to the host while the next computes.
// 32768^3: 365 ms (row chunks only) -> ~300 ms; the device-resident step is 273 ms.
static int host_gemm_f32_kpipelined(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t rsA, const float* B,
                                    int64_t rsB, float* C, int64_t rsC) {
  const int NS = 4;                                          // K slices of phase 1
  int dev = 0;
  AM_CUDA_TRY(cudaGetDevice(&dev));
  cudaMemPool_t pool;
  int status = host_pool(dev, &pool);
  if (status) return status;
  HostCtx& hc = t_host;
  if ((status = hc.init(dev))) return status;
  cudaStream_t s_in = hc.s_in, s_cmp = hc.s_cmp, s_out = hc.s_out;
  const int64_t K1 = (K / 2) / (32 * NS) * (32 * NS), kc = K1 / NS, K2 = K - K1;
  const int64_t rows = ((M / 8 + 255) / 256) * 256;           // row chunk of phase 2
  const int nrc = (int)((M + rows - 1) / rows);
  const int64_t last_rows = M - (int64_t)(nrc - 1) * rows;
  float *dA = nullptr, *dB = nullptr, *dC = nullptr, *pk = nullptr;
  void *hA1 = nullptr, *hB1 = nullptr, *hB2 = nullptr, *hA2 = nullptr, *hA2l = nullptr;
  cudaError_t e = cudaSuccess;
  const int64_t fA1 = packed_floats_f32(M, kc), fB1 = packed_floats_f32(N, kc), fB2 = packed_floats_f32(N, K2),
                fA2 = packed_floats_f32(rows, K2);
  do {
    if ((e = cudaMallocFromPoolAsync((void**)&dA, (size_t)(M * K) * 4, pool, s_in)) != cudaSuccess ||
        (e = cudaMallocFromPoolAsync((void**)&dB, (size_t)(K * N) * 4, pool, s_in)) != cudaSuccess ||
        (e = cudaMallocFromPoolAsync((void**)&dC, (size_t)(M * N) * 4, pool, s_in)) != cudaSuccess ||
        (e = cudaMallocFromPoolAsync((void**)&pk, (size_t)(fA1 + fB1 + fB2 + fA2) * 4, pool, s_in)) != cudaSuccess) { status = cuda_fail(e, "cudaMallocFromPoolAsync"); break; }
    float* pA1 = pk; float* pB1 = pA1 + fA1; float* pB2 = pB1 + fB1; float* pA2 = pB2 + fB2;
    std::vector<cudaEvent_t> ev1((size_t)NS), evA2((size_t)nrc), evC((size_t)nrc);
    cudaEvent_t ev_alloc = hc.event(), evB2 = hc.event();
    for (auto& ev : ev1) ev = hc.event();
    for (auto& ev : evA2) ev = hc.event();
    for (auto& ev : evC) ev = hc.event();
    if (!ev_alloc || !evB2 || !ev1.back() || !evA2.back() || !evC.back()) { status = cuda_fail(cudaErrorMemoryAllocation, "event"); break; }
    cudaEventRecord(ev_alloc, s_in);
    cudaStreamWaitEvent(s_cmp, ev_alloc, 0);
    cudaStreamWaitEvent(s_out, ev_alloc, 0);
    // ---- uploads, in the order the products need them (one stream: the PCIe link is the shared resource)
    for (int c = 0; c < NS && !status; c++) {
      const int64_t k0 = c * kc;
      if ((e = cudaMemcpy2DAsync(dB + k0 * N, (size_t)N * 4, B + k0 * rsB, (size_t)rsB * 4, (size_t)N * 4, (size_t)kc, cudaMemcpyHostToDevice, s_in)) != cudaSuccess ||
          (e = cudaMemcpy2DAsync(dA + k0, (size_t)K * 4, A + k0, (size_t)rsA * 4, (size_t)kc * 4, (size_t)M, cudaMemcpyHostToDevice, s_in)) != cudaSuccess) { status = cuda_fail(e, "H2D K slice"); break; }
      cudaEventRecord(ev1[c], s_in);
    }
    if (status) break;
    if ((e = cudaMemcpy2DAsync(dB + K1 * N, (size_t)N * 4, B + K1 * rsB, (size_t)rsB * 4, (size_t)N * 4, (size_t)K2, cudaMemcpyHostToDevice, s_in)) != cudaSuccess) { status = cuda_fail(e, "H2D B tail"); break; }
    cudaEventRecord(evB2, s_in);
    for (int i = 0; i < nrc && !status; i++) {
      const int64_t r0 = (int64_t)i * rows, nr = (i == nrc - 1) ? last_rows : rows;
      if ((e = cudaMemcpy2DAsync(dA + r0 * K + K1, (size_t)K * 4, A + r0 * rsA + K1, (size_t)rsA * 4, (size_t)K2 * 4, (size_t)nr, cudaMemcpyHostToDevice, s_in)) != cudaSuccess) { status = cuda_fail(e, "H2D A tail"); break; }
      cudaEventRecord(evA2[i], s_in);
    }
    if (status) break;
    // ---- phase 1: K slices over the whole C
    for (int c = 0; c < NS && !status; c++) {
      const int64_t k0 = c * kc;
      cudaStreamWaitEvent(s_cmp, ev1[c], 0);
      if (c == 0) {
        if ((status = pack_f32_view(s_cmp, M, kc, dA + k0, K, 1, pA1, &hA1))) break;
        if ((status = pack_f32_view(s_cmp, N, kc, dB + k0 * N, 1, N, pB1, &hB1))) break;
      } else {
        if ((status = repack_f32(s_cmp, hA1, dA + k0, K, 1))) break;
        if ((status = repack_f32(s_cmp, hB1, dB + k0 * N, 1, N))) break;
      }
      status = gemm_packed_f32(s_cmp, alpha, hA1, hB1, c == 0 ? 0.f : 1.f, dC, N, 1);
    }
    if (status) break;
    // ---- phase 2: the rest of K by row chunks, results streaming back
    cudaStreamWaitEvent(s_cmp, evB2, 0);
    if ((status = pack_f32_view(s_cmp, N, K2, dB + K1 * N, 1, N, pB2, &hB2))) break;
    for (int i = 0; i < nrc && !status; i++) {
      const int64_t r0 = (int64_t)i * rows, nr = (i == nrc - 1) ? last_rows : rows;
      cudaStreamWaitEvent(s_cmp, evA2[i], 0);
      void** hh = (nr == rows) ? &hA2 : &hA2l;
      if (*hh == nullptr) status = pack_f32_view(s_cmp, nr, K2, dA + r0 * K + K1, K, 1, pA2, hh);
      else status = repack_f32(s_cmp, *hh, dA + r0 * K + K1, K, 1);
      if (status) break;
      if ((status = gemm_packed_f32(s_cmp, alpha, *hh, hB2, 1.f, dC + r0 * N, N, 1))) break;
      cudaEventRecord(evC[i], s_cmp);
      cudaStreamWaitEvent(s_out, evC[i], 0);
      if ((e = cudaMemcpy2DAsync(C + r0 * rsC, (size_t)rsC * 4, dC + r0 * N, (size_t)N * 4, (size_t)N * 4, (size_t)nr, cudaMemcpyDeviceToHost, s_out)) != cudaSuccess) { status = cuda_fail(e, "D2H C chunk"); break; }
    }
  } while (0);
  if ((e = cudaStreamSynchronize(s_in)) != cudaSuccess && !status) status = cuda_fail(e, "sync");
  if ((e = cudaStreamSynchronize(s_cmp)) != cudaSuccess && !status) status = cuda_fail(e, "sync");
  if ((e = cudaStreamSynchronize(s_out)) != cudaSuccess && !status) status = cuda_fail(e, "sync");
  for (void* h : {hA1, hB1, hB2, hA2, hA2l}) if (h) packed_free_f32(h);
  if (dA) cudaFreeAsync(dA, s_in);
  if (dB) cudaFreeAsync(dB, s_in);
  if (dC) cudaFreeAsync(dC, s_in);
  if (pk) cudaFreeAsync(pk, s_in);
  cudaStreamSynchronize(s_in);
  return status;
}

// host-buffer GEMM: what `a.cuda * b.cuda` then `.cpu` does (init_cuda.nim:23-59), as one call.
// Row-major-like A and C are processed in row chunks on three streams — H2D of chunk j+1, GEMM of chunk j and
// D2H of chunk j-1 overlap — after B has been copied once.  `device_gemm(stream, chunk_index, ...)` is the per-chunk
// product; the float32 functor splits/packs B once (chunk 0) into an explicit handle and reuses it for the other chunks.
template <class T, class F>
static int host_gemm(F&& device_gemm, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t rsA,
                     int64_t csA, const T* B, int64_t rsB, int64_t csB, T beta, T* C, int64_t rsC, int64_t csC) {
  int rc = check_gemm_args(M, N, K, A, B, C);
  if (rc == -1) return AM_OK;
  if (rc) return rc;
  // Bounding extents of each (possibly negatively strided) view, in elements.
  auto extent = [](int64_t r, int64_t c, int64_t rs, int64_t cs, int64_t* lo, int64_t* hi) {
    int64_t a = (r - 1) * rs, b = (c - 1) * cs;
    *lo = (a < 0 ? a : 0) + (b < 0 ? b : 0);
    *hi = (a > 0 ? a : 0) + (b > 0 ? b : 0);
  };
  int64_t loA, hiA, loB, hiB, loC, hiC;
  extent(M, K, rsA, csA, &loA, &hiA);
  extent(K, N, rsB, csB, &loB, &hiB);
  extent(M, N, rsC, csC, &loC, &hiC);
  const size_t nA = (size_t)(hiA - loA + 1), nB = (size_t)(hiB - loB + 1), nC = (size_t)(hiC - loC + 1);

  int dev = 0;
  AM_CUDA_TRY(cudaGetDevice(&dev));
  cudaMemPool_t pool;
  int status = host_pool(dev, &pool);
  if (status) return status;
  HostCtx& hc = t_host;
  if ((status = hc.init(dev))) return status;
  cudaStream_t s_in = hc.s_in, s_cmp = hc.s_cmp, s_out = hc.s_out;
  // row chunks: only when rows of A and of C are increasing, disjoint host ranges and C's rows are dense runs of N
  // elements (csC == 1): each chunk of C then goes back with ONE 2-D copy that touches exactly the N valid elements of
  // every row, so whatever the host holds between the rows (rsC > N) is left alone.
  const bool rowwise = rsA > 0 && csA > 0 && rsA >= (K - 1) * csA + 1 && csC == 1 && rsC >= N && beta == T(0);
  int64_t chunk = M;
  if (rowwise && M >= 2048) {
    chunk = ((M / 8 + 255) / 256) * 256;
    if (chunk < 512) chunk = 512;
  }
  const int nchunks = (int)((M + chunk - 1) / chunk);

  T *dA = nullptr, *dB = nullptr, *dC = nullptr;
  cudaError_t e = cudaSuccess;
  do {
    if ((e = cudaMallocFromPoolAsync((void**)&dA, nA * sizeof(T), pool, s_in)) != cudaSuccess ||
        (e = cudaMallocFromPoolAsync((void**)&dB, nB * sizeof(T), pool, s_in)) != cudaSuccess ||
        (e = cudaMallocFromPoolAsync((void**)&dC, nC * sizeof(T), pool, s_in)) != cudaSuccess) { status = cuda_fail(e, "cudaMallocFromPoolAsync"); break; }
    if ((e = cudaMemcpyAsync(dB, B + loB, nB * sizeof(T), cudaMemcpyHostToDevice, s_in)) != cudaSuccess) { status = cuda_fail(e, "H2D B"); break; }
    if (nchunks == 1) {
      if ((e = cudaMemcpyAsync(dA, A + loA, nA * sizeof(T), cudaMemcpyHostToDevice, s_in)) != cudaSuccess) { status = cuda_fail(e, "H2D A"); break; }
      if (beta != T(0) || nC != (size_t)(M * N)) {   // C is read, or the view has gaps that must survive the D2H copy
        if ((e = cudaMemcpyAsync(dC, C + loC, nC * sizeof(T), cudaMemcpyHostToDevice, s_in)) != cudaSuccess) { status = cuda_fail(e, "H2D C"); break; }
      }
      status = device_gemm(s_in, 0, M, N, K, alpha, dA - loA, rsA, csA, dB - loB, rsB, csB, beta, dC - loC, rsC, csC);
      if (status) break;
      if ((e = cudaMemcpyAsync(C + loC, dC, nC * sizeof(T), cudaMemcpyDeviceToHost, s_in)) != cudaSuccess) { status = cuda_fail(e, "D2H"); break; }
    } else {
      cudaEvent_t ev_alloc = hc.event();
      if (!ev_alloc) { status = cuda_fail(cudaErrorMemoryAllocation, "event"); break; }
      cudaEventRecord(ev_alloc, s_in);                 // allocations (and the B copy) are ordered on s_in
      cudaStreamWaitEvent(s_out, ev_alloc, 0);
      for (int j = 0; j < nchunks && !status; j++) {
        const int64_t r0 = (int64_t)j * chunk, rows = (M - r0 < chunk) ? M - r0 : chunk;
        const size_t a_elems = (size_t)((rows - 1) * rsA + (K - 1) * csA + 1);
        cudaEvent_t ev_in = hc.event(), ev_cmp = hc.event();
        if (!ev_in || !ev_cmp) { status = cuda_fail(cudaErrorMemoryAllocation, "event"); break; }
        if ((e = cudaMemcpyAsync(dA + r0 * rsA, A + r0 * rsA, a_elems * sizeof(T), cudaMemcpyHostToDevice, s_in)) != cudaSuccess) { status = cuda_fail(e, "H2D A chunk"); break; }
        cudaEventRecord(ev_in, s_in);
        cudaStreamWaitEvent(s_cmp, ev_in, 0);          // (also orders after the B copy, issued earlier on s_in)
        status = device_gemm(s_cmp, j, rows, N, K, alpha, dA + r0 * rsA, rsA, csA, dB - loB, rsB, csB, beta, dC + r0 * rsC, rsC, csC);
        if (status) break;
        cudaEventRecord(ev_cmp, s_cmp);
        cudaStreamWaitEvent(s_out, ev_cmp, 0);
        // exactly the N valid elements of each row (pitch rsC): gaps between the rows of a padded host C are not written
        if ((e = cudaMemcpy2DAsync(C + r0 * rsC, (size_t)rsC * sizeof(T), dC + r0 * rsC, (size_t)rsC * sizeof(T), (size_t)N * sizeof(T),
                                   (size_t)rows, cudaMemcpyDeviceToHost, s_out)) != cudaSuccess) { status = cuda_fail(e, "D2H C chunk"); break; }
      }
    }
  } while (0);
  if ((e = cudaStreamSynchronize(s_in)) != cudaSuccess && !status) status = cuda_fail(e, "sync");
  if ((e = cudaStreamSynchronize(s_cmp)) != cudaSuccess && !status) status = cuda_fail(e, "sync");
  if ((e = cudaStreamSynchronize(s_out)) != cudaSuccess && !status) status = cuda_fail(e, "sync");
  if (dA) cudaFreeAsync(dA, s_in);
  if (dB) cudaFreeAsync(dB, s_in);
  if (dC) cudaFreeAsync(dC, s_in);
  cudaStreamSynchronize(s_in);
  return status;
}

}  // namespace am

using namespace am;

extern "C" {

const char* am_version(void) { return "arraymancer_b200 0.1 (sm_100a)"; }
const char* am_last_error(void) { return g_err; }

int am_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  AM_CUDA_TRY(cudaGetDevice(&dev));
  int a = 0, b = 0, c = 0;
  AM_CUDA_TRY(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  AM_CUDA_TRY(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  AM_CUDA_TRY(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sms) *sms = a;
  if (major) *major = b;
  if (minor) *minor = c;
  return AM_OK;
}

int am_shutdown(void) {
  workspace_release_all();
  host_pools_release();          // the private stream-ordered pools of the host-buffer entries go back to the driver
  return AM_OK;
}

int am_set_tuning(const char* name, int value) {
  const int k = tune_key(name);
  if (k < 0) { set_last_error("am_set_tuning: unknown knob '%s'", name ? name : "(null)"); return AM_ERR_INVALID; }
  g_tune[k].store(value);
  return AM_OK;
}
int am_get_tuning(const char* name, int* value) {
  const int k = tune_key(name);
  if (k < 0 || !value) { set_last_error("am_get_tuning: unknown knob '%s'", name ? name : "(null)"); return AM_ERR_INVALID; }
  *value = g_tune[k].load();
  return AM_OK;
}

int am_set_f32_path(int path) {
  if (path < AM_F32_AUTO || path > AM_F32_TC_1CTA) { set_last_error("am_set_f32_path: bad selector"); return AM_ERR_INVALID; }
  g_f32_path.store(path);
  return AM_OK;
}
int am_get_f32_path(void) { return g_f32_path.load(); }

int64_t am_kernel_launch_count(void) { return g_launch_count.load(); }
int am_microbench(int which, double* tops) { return microbench(which, tops); }

int am_gemm_strided_f32(am_stream_t s, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t rsA,
                        int64_t csA, const float* B, int64_t rsB, int64_t csB, float beta, float* C, int64_t rsC,
                        int64_t csC) {
  int rc = check_gemm_args(M, N, K, A, B, C);
  if (rc == -1) return AM_OK;
  if (rc) return rc;
  return gemm_dispatch_f32((cudaStream_t)s, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, nullptr);
}

#define DEF_GEMM_SIMT(SUF, T)                                                                                  \
  int am_gemm_strided_##SUF(am_stream_t s, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t rsA,   \
                            int64_t csA, const T* B, int64_t rsB, int64_t csB, T beta, T* C, int64_t rsC,       \
                            int64_t csC) {                                                                      \
    int rc = check_gemm_args(M, N, K, A, B, C);                                                                 \
    if (rc == -1) return AM_OK;                                                                                 \
    if (rc) return rc;                                                                                          \
    bool done = false;                                                                                          \
    rc = gemm_skinny<T>((cudaStream_t)s, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, &done);   \
    if (rc || done) return rc;                                                                                  \
    return gemm_simt<T>((cudaStream_t)s, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC);          \
  }
int am_set_f64_path(int path) {
  if (path < AM_F64_AUTO || path > AM_F64_DMMA) { set_last_error("am_set_f64_path: bad selector"); return AM_ERR_INVALID; }
  g_f64_path.store(path);
  return AM_OK;
}
int am_get_f64_path(void) { return g_f64_path.load(); }

int am_gemm_strided_f64(am_stream_t s, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t rsA,
                        int64_t csA, const double* B, int64_t rsB, int64_t csB, double beta, double* C, int64_t rsC,
                        int64_t csC) {
  int rc = check_gemm_args(M, N, K, A, B, C);
  if (rc == -1) return AM_OK;
  if (rc) return rc;
  return gemm_dispatch_f64((cudaStream_t)s, M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, nullptr);
}
DEF_GEMM_SIMT(i32, int32_t)
DEF_GEMM_SIMT(i64, int64_t)

int am_pack_f32_a(am_stream_t s, int64_t M, int64_t K, const float* A, int64_t rsA, int64_t csA, am_packed_f32** out) {
  return pack_f32((cudaStream_t)s, M, K, A, rsA, csA, (void**)out);
}
int am_pack_f32_b(am_stream_t s, int64_t K, int64_t N, const float* B, int64_t rsB, int64_t csB, am_packed_f32** out) {
  return pack_f32((cudaStream_t)s, N, K, B, csB, rsB, (void**)out);
}
int am_repack_f32_a(am_stream_t s, am_packed_f32* h, const float* A, int64_t rsA, int64_t csA) {
  return repack_f32((cudaStream_t)s, h, A, rsA, csA);
}
int am_repack_f32_b(am_stream_t s, am_packed_f32* h, const float* B, int64_t rsB, int64_t csB) {
  return repack_f32((cudaStream_t)s, h, B, csB, rsB);
}
int am_gemm_packed_f32(am_stream_t s, float alpha, const am_packed_f32* A, const am_packed_f32* B, float beta,
                       float* C, int64_t rsC, int64_t csC) {
  return gemm_packed_f32((cudaStream_t)s, alpha, A, B, beta, C, rsC, csC);
}
int am_gemm_packed_f32_bcast(am_stream_t s, float alpha, const am_packed_f32* A, const am_packed_f32* B, int npeers,
                             float* const* peerC, int self_index, int64_t rsC, int64_t csC) {
  return gemm_packed_f32_bcast((cudaStream_t)s, alpha, A, B, npeers, peerC, self_index, rsC, csC);
}
int am_packed_free_f32(am_packed_f32* h) { return packed_free_f32(h); }

// cublas_gemm adapter (cublas.nim:142-170): column-major, op N -> (rs=1, cs=ld), op T -> (rs=ld, cs=1)
#define DEF_CUBLAS(SUF, T)                                                                                     \
  int am_cublas_gemm_##SUF(am_stream_t s, int transa, int transb, int64_t m, int64_t n, int64_t k, T alpha,     \
                           const T* A, int64_t lda, const T* B, int64_t ldb, T beta, T* C, int64_t ldc) {       \
    if ((transa != 0 && transa != 1) || (transb != 0 && transb != 1)) {                                         \
      set_last_error("cublas_gemm: only CUBLAS_OP_N (0) / CUBLAS_OP_T (1) are supported");                      \
      return AM_ERR_INVALID;                                                                                    \
    }                                                                                                           \
    const int64_t rowsA = transa ? k : m, rowsB = transb ? n : k;                                               \
    if (lda < (rowsA > 1 ? rowsA : 1) || ldb < (rowsB > 1 ? rowsB : 1) || ldc < (m > 1 ? m : 1)) {              \
      set_last_error("cublas_gemm: leading dimension smaller than the matrix rows");                            \
      return AM_ERR_NONCONTIGUOUS;                                                                              \
    }                                                                                                           \
    return am_gemm_strided_##SUF(s, m, n, k, alpha, A, transa ? lda : 1, transa ? 1 : lda, B, transb ? ldb : 1, \
                                 transb ? 1 : ldb, beta, C, 1, ldc);                                            \
  }
DEF_CUBLAS(f32, float)
DEF_CUBLAS(f64, double)

int am_set_conv_path(int path) {
  if (path < AM_CONV_AUTO || path > AM_CONV_TC) { set_last_error("am_set_conv_path: bad selector"); return AM_ERR_INVALID; }
  am::g_conv_path.store(path);
  return AM_OK;
}

int am_conv2d_out_dims(const am_conv2d_desc* d, int64_t* Ho, int64_t* Wo) {
  if (!d || d->strideH <= 0 || d->strideW <= 0 || d->dilH <= 0 || d->dilW <= 0) { set_last_error("conv2d_out_dims: bad descriptor"); return AM_ERR_INVALID; }
  if (Ho) *Ho = (d->H + 2 * d->padH - (d->dilH * (d->kH - 1) + 1)) / d->strideH + 1;
  if (Wo) *Wo = (d->W + 2 * d->padW - (d->dilW * (d->kW - 1) + 1)) / d->strideW + 1;
  return AM_OK;
}

#define DEF_CONV(SUF, T)                                                                                       \
  int am_conv2d_forward_##SUF(am_stream_t s, const am_conv2d_desc* d, const T* in, const T* k, const T* b,      \
                              T* out) {                                                                         \
    if (!d) { set_last_error("conv2d_forward: null descriptor"); return AM_ERR_INVALID; }                      \
    return conv2d_forward<T>((cudaStream_t)s, *d, in, k, b, out);                                               \
  }                                                                                                             \
  int am_conv2d_forward_act_##SUF(am_stream_t s, const am_conv2d_desc* d, const T* in, const T* k, const T* b,  \
                                  T* out, int activation) {                                                     \
    if (!d) { set_last_error("conv2d_forward: null descriptor"); return AM_ERR_INVALID; }                      \
    if (activation != AM_ACT_NONE && activation != AM_ACT_RELU) {                                               \
      set_last_error("conv2d_forward_act: unknown activation %d", activation); return AM_ERR_INVALID;           \
    }                                                                                                           \
    return conv2d_forward<T>((cudaStream_t)s, *d, in, k, b, out, activation);                                   \
  }                                                                                                             \
  int am_conv2d_backward_##SUF(am_stream_t s, const am_conv2d_desc* d, const T* in, const T* k, const T* go,    \
                               T* gi, T* gk, T* gb) {                                                           \
    if (!d) { set_last_error("conv2d_backward: null descriptor"); return AM_ERR_INVALID; }                     \
    return conv2d_backward<T>((cudaStream_t)s, *d, in, k, go, gi, gk, gb);                                      \
  }
DEF_CONV(f32, float)
DEF_CONV(f64, double)
DEF_CONV(i32, int32_t)
DEF_CONV(i64, int64_t)

#define DEF_NN(SUF, T)                                                                                         \
  int am_relu_forward_##SUF(am_stream_t s, int64_t n, const T* x, T* y) { return relu_forward<T>((cudaStream_t)s, n, x, y); } \
  int am_relu_backward_##SUF(am_stream_t s, int64_t n, const T* g, const T* c, T* o) {                          \
    return relu_backward<T>((cudaStream_t)s, n, g, c, o);                                                      \
  }                                                                                                            \
  int am_maxpool2d_forward_##SUF(am_stream_t s, int64_t N, int64_t C, int64_t H, int64_t W, int64_t kH, int64_t kW, \
                                 int64_t pH, int64_t pW, int64_t sH, int64_t sW, const T* x, T* y, int64_t* idx) { \
    return maxpool2d_forward<T>((cudaStream_t)s, N, C, H, W, kH, kW, pH, pW, sH, sW, x, y, idx);               \
  }                                                                                                            \
  int am_maxpool2d_backward_##SUF(am_stream_t s, int64_t n_in, int64_t n_out, const int64_t* idx, const T* go, T* gi, \
                                  int overlap) {                                                               \
    return maxpool2d_backward<T>((cudaStream_t)s, n_in, n_out, idx, go, nullptr, gi, overlap);                 \
  }                                                                                                            \
  int am_maxpool2d_backward_relu_##SUF(am_stream_t s, int64_t n_in, int64_t n_out, const int64_t* idx, const T* go, \
                                       const T* relu_cached, T* gi, int overlap) {                             \
    if (n_in > 0 && !relu_cached) { set_last_error("maxpool2d_backward_relu: null cached tensor"); return AM_ERR_INVALID; } \
    return maxpool2d_backward<T>((cudaStream_t)s, n_in, n_out, idx, go, relu_cached, gi, overlap);             \
  }                                                                                                            \
  int am_linear_forward_##SUF(am_stream_t s, int64_t b, int64_t in, int64_t out, const T* x, const T* w, const T* bias, T* y) { \
    return linear_forward<T>((cudaStream_t)s, b, in, out, x, w, bias, y);                                      \
  }                                                                                                            \
  int am_linear_backward_##SUF(am_stream_t s, int64_t b, int64_t in, int64_t out, const T* x, const T* w, const T* go, \
                               T* gi, T* gw, T* gb) {                                                          \
    return linear_backward<T>((cudaStream_t)s, b, in, out, x, w, go, gi, gw, gb);                              \
  }                                                                                                            \
  int am_sparse_softmax_cross_entropy_##SUF(am_stream_t s, int64_t b, int64_t f, const T* x, int64_t rs, int64_t cs, \
                                            const int64_t* labels, T* loss) {                                  \
    return ssce_forward<T>((cudaStream_t)s, b, f, x, rs, cs, labels, loss);                                    \
  }                                                                                                            \
  int am_sparse_softmax_cross_entropy_backward_##SUF(am_stream_t s, int64_t b, int64_t f, T grad, const T* x, int64_t rs, \
                                                     int64_t cs, const int64_t* labels, T* out) {              \
    return ssce_backward<T>((cudaStream_t)s, b, f, grad, x, rs, cs, labels, out);                              \
  }
DEF_NN(f32, float)
DEF_NN(f64, double)

#define DEF_HOST(SUF, T)                                                                                       \
  int am_host_gemm_strided_##SUF(int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t rsA, int64_t csA, \
                                 const T* B, int64_t rsB, int64_t csB, T beta, T* C, int64_t rsC, int64_t csC) { \
    return host_gemm<T>(                                                                                        \
        [](cudaStream_t st, int, int64_t m, int64_t n, int64_t k, T al, const T* a, int64_t ra, int64_t ca,     \
           const T* b, int64_t rb, int64_t cb, T be, T* c, int64_t rc_, int64_t cc) {                           \
          return am_gemm_strided_##SUF((am_stream_t)st, m, n, k, al, a, ra, ca, b, rb, cb, be, c, rc_, cc);     \
        },                                                                                                      \
        M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC);                                           \
  }
int am_host_gemm_strided_f32(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t rsA, int64_t csA,
                             const float* B, int64_t rsB, int64_t csB, float beta, float* C, int64_t rsC, int64_t csC) {
  // large row-major products on the tensor cores: K-pipelined uploads (see host_gemm_f32_kpipelined)
  const int path = g_f32_path.load();
  if (!tuning(kTuneHostRowChunks) && (path == AM_F32_AUTO || path == AM_F32_TC) && A && B && C && beta == 0.f && csA == 1 && csB == 1 && csC == 1 &&
      rsA >= K && rsB >= N && rsC >= N && M >= 4096 && N >= 4096 && K >= 4096 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31) &&
      gemm_f32_tc_available())
    return host_gemm_f32_kpipelined(M, N, K, alpha, A, rsA, B, rsB, C, rsC);
  // row chunks: B is split/packed ONCE into an explicit handle by the first chunk that takes the tensor-core path and
  // handed to the following chunks (no hidden "previous B" state in the GEMM)
  void* hB = nullptr;
  const bool tc_whole = (path == AM_F32_TC) || (path == AM_F32_AUTO && gemm_f32_tc_available() && M >= 256 && N >= 256 &&
                                                K >= 256 && 2.0 * (double)M * (double)N * (double)K >= 3.0e9);
  int rc = host_gemm<float>(
      [&](cudaStream_t st, int, int64_t m, int64_t n, int64_t k, float al, const float* a, int64_t ra, int64_t ca, const float* b,
          int64_t rb, int64_t cb, float be, float* c, int64_t rc_, int64_t cc) {
        if (tc_whole && m >= 256 && gemm_f32_tc_available()) {
          if (!hB) { int r = pack_f32(st, n, k, b, cb, rb, &hB); if (r) return r; }
          return gemm_f32_tc(st, 2, m, n, k, al, a, ra, ca, b, rb, cb, be, c, rc_, cc, hB);
        }
        return am_gemm_strided_f32((am_stream_t)st, m, n, k, al, a, ra, ca, b, rb, cb, be, c, rc_, cc);
      },
      M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC);
  if (hB) packed_free_f32(hB);
  return rc;
}
DEF_HOST(f64, double)
DEF_HOST(i32, int32_t)
DEF_HOST(i64, int64_t)

}  // extern "C"
