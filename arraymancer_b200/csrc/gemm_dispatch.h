// Internal declarations shared by the translation units of libarraymancer_b200.so.
#pragma once
#include <atomic>
#include <cmath>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/am_b200.h"

namespace am {

extern std::atomic<int64_t> g_launch_count;

// Explicit tuning knobs (am_set_tuning / am_get_tuning in the C ABI): process-wide atomics set by the caller; the
// library reads NO environment variables.
enum TuneKey : int {
  kTuneTcFlushKb = 0,      // "tc_flush_kb": k-blocks (of 32) per tensor-core accumulation chain (default 2)
  kTuneTcGroup,            // "tc_group": rasterisation group of the persistent GEMM (default 8; < 0: groups of N tiles)
  kTuneTcSync,             // "tc_sync": per-wave grid barrier of the persistent GEMM (default 1)
  kTunePackScalar,         // "pack_scalar": force the scalar split/pack kernel (default 0)
  kTuneHostRowChunks,      // "host_rowchunks": host-buffer f32 GEMM uses row chunks only, no K pipeline (default 0)
  kTuneConvTcGroups,       // "convtc_groups": gather-warp groups of the tcgen05 conv kernels (0 = kernel default)
  kTuneConvTcDebug,        // "convtc_debug": per-role wait-cycle counters of the tcgen05 conv kernels (default 0)
  kTuneConvDgradGather,    // "convtc_dgrad_gather": older gather-form tcgen05 dgrad kernel under AM_CONV_TC (default 0)
  kTuneSimtVecLoad,        // "simt_vec_load": 128-bit global loads along a unit-stride operand dimension in the SIMT GEMM (default 1)
  kTuneDmmaTma,            // "dmma_tma": TMA-fed 16-warp DMMA kernel for unit-stride float64 operands (default 1; 0: register-staged kernel)
  kTuneConvTcHiRes,        // "convtc_hi_resident": tcgen05 conv forward keeps the hi weight planes resident in shared memory (default 1)
  kTuneConvTcFlushKb,      // "convtc_flush_kb": k blocks (of 32) per TMEM accumulation chain of the tcgen05 conv forward (default 4)
  kTuneConvTcWgradTma,     // "convtc_wgrad_tma": tcgen05 wgrad fetches grad_output stages by TMA from pre-split hi / lo planes (default 1)
  kTuneCount
};
int tuning(int key);
extern std::atomic<int> g_conv_path;          // conv.cu: AM_CONV_AUTO / AM_CONV_GATHER

// gemm_simt.cu — strided SIMT GEMM, any layout
template <class T>
int gemm_simt(cudaStream_t st, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t rsA, int64_t csA,
              const T* B, int64_t rsB, int64_t csB, T beta, T* C, int64_t rsC, int64_t csC, const T* bias_col = nullptr);
// gemm_skinny.cu — DRAM-bound products with min(M, N) <= 16; *done == false -> not handled, use the general kernels
template <class T>
int gemm_skinny(cudaStream_t st, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t rsA, int64_t csA, const T* B,
                int64_t rsB, int64_t csB, T beta, T* C, int64_t rsC, int64_t csC, bool* done);
// am_api.cu — the dispatch behind am_gemm_strided_f32 / _f64, with the optional fused per-column bias of the linear layer
int gemm_dispatch_f32(cudaStream_t st, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t rsA, int64_t csA,
                      const float* B, int64_t rsB, int64_t csB, float beta, float* C, int64_t rsC, int64_t csC, const float* bias_col);
int gemm_dispatch_f64(cudaStream_t st, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t rsA, int64_t csA,
                      const double* B, int64_t rsB, int64_t csB, double beta, double* C, int64_t rsC, int64_t csC, const double* bias_col);

// one launch for `batch` independent products (operand b at X + b*bsX), blockIdx.z = b
template <class T>
int gemm_simt_batched(cudaStream_t st, int64_t batch, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t rsA,
                      int64_t csA, int64_t bsA, const T* B, int64_t rsB, int64_t csB, int64_t bsB, T beta, T* C,
                      int64_t rsC, int64_t csC, int64_t bsC);

// gemm_f32_tc.cu — tcgen05 3xTF32 GEMM (split/pack pre-pass + TMA/UMMA mainloop)
// cta_group: 1 or 2
int gemm_f32_tc(cudaStream_t st, int cta_group, int64_t M, int64_t N, int64_t K, float alpha, const float* A,
                int64_t rsA, int64_t csA, const float* B, int64_t rsB, int64_t csB, float beta, float* C,
                int64_t rsC, int64_t csC, const void* prepackedB = nullptr, const float* bias_col = nullptr);
bool gemm_f32_tc_available();
// pre-packed operands (two tf32 planes, K-major): pack once, multiply many
int pack_f32(cudaStream_t st, int64_t R, int64_t K, const float* X, int64_t r_stride, int64_t k_stride, void** handle);
int repack_f32(cudaStream_t st, void* handle, const float* X, int64_t r_stride, int64_t k_stride);
int64_t packed_floats_f32(int64_t R, int64_t K);
int pack_f32_view(cudaStream_t st, int64_t R, int64_t K, const float* X, int64_t r_stride, int64_t k_stride, float* planes,
                  void** handle);
int packed_wrap_f32(int64_t R, int64_t K, float* planes, void** handle);
int packed_free_f32(void* handle);
int gemm_packed_f32_bcast(cudaStream_t st, float alpha, const void* hA, const void* hB, int npeers, float* const* peers,
                          int self, int64_t rsC, int64_t csC);
int gemm_packed_f32(cudaStream_t st, float alpha, const void* hA, const void* hB, float beta, float* C, int64_t rsC,
                    int64_t csC);

// gemm_f64_dmma.cu — FP64 tensor-pipe GEMM (mma.sync m8n8k4 f64), any strides
int gemm_f64_dmma(cudaStream_t st, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t rsA,
                  int64_t csA, const double* B, int64_t rsB, int64_t csB, double beta, double* C, int64_t rsC,
                  int64_t csC, const double* bias_col = nullptr);

// conv.cu
template <class T>
int conv2d_forward(cudaStream_t st, const am_conv2d_desc& d, const T* input, const T* kernel, const T* bias,
                   T* output, int act = 0);      // act: AM_ACT_NONE / AM_ACT_RELU fused into the epilogue
template <class T>
int conv2d_backward(cudaStream_t st, const am_conv2d_desc& d, const T* input, const T* kernel,
                    const T* grad_output, T* grad_input, T* grad_kernel, T* grad_bias);

// nn_ops.cu — ReLU, MaxPool2D, linear-layer bias passes, sparse softmax cross-entropy
template <class T> int relu_forward(cudaStream_t st, int64_t n, const T* x, T* y);
template <class T> int relu_backward(cudaStream_t st, int64_t n, const T* grad, const T* cached, T* out);
template <class T>
int maxpool2d_forward(cudaStream_t st, int64_t N, int64_t C, int64_t H, int64_t W, int64_t kH, int64_t kW, int64_t padH,
                      int64_t padW, int64_t sH, int64_t sW, const T* x, T* y, int64_t* idx);
template <class T>
int maxpool2d_backward(cudaStream_t st, int64_t n_in, int64_t n_out, const int64_t* idx, const T* go, const T* relu_cached, T* gi,
                       int windows_overlap);
template <class T>
int linear_forward(cudaStream_t st, int64_t batch, int64_t in, int64_t out, const T* x, const T* w, const T* bias, T* y);
template <class T>
int linear_backward(cudaStream_t st, int64_t batch, int64_t in, int64_t out, const T* x, const T* w, const T* go, T* gi,
                    T* gw, T* gb);
template <class T>
int ssce_forward(cudaStream_t st, int64_t batch, int64_t features, const T* x, int64_t rs, int64_t cs, const int64_t* labels,
                 T* loss_dev);
template <class T>
int ssce_backward(cudaStream_t st, int64_t batch, int64_t features, T grad, const T* x, int64_t rs, int64_t cs,
                  const int64_t* labels, T* out);

// peaks.cu
int microbench(int which, double* tops);

}  // namespace am
