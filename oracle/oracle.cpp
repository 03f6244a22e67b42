// ORACLE — TEST INFRASTRUCTURE ONLY.  C ABI over the laser / im2col restatements.
// See laser_gemm.hpp and conv_oracle.hpp for the reference file:line each routine follows.
// Built by oracle/Makefile into oracle/liblaser_oracle.so (x86-64-v3: AVX2+FMA, the
// reference's default build tops out at AVX+FMA — SURVEY F9).
#include "conv_oracle.hpp"

using namespace laser_oracle;

#define EXPORT extern "C" __attribute__((visibility("default")))

#define DEF_GEMM(SUF, T)                                                                               \
  EXPORT void oracle_gemm_strided_##SUF(int64_t M, int64_t N, int64_t K, T alpha, const T* A,          \
                                        int64_t rsA, int64_t csA, const T* B, int64_t rsB,             \
                                        int64_t csB, T beta, T* C, int64_t rsC, int64_t csC,           \
                                        int variant, int threads) {                                    \
    gemm_strided<T>(M, N, K, alpha, A, rsA, csA, B, rsB, csB, beta, C, rsC, csC, variant, threads);    \
  }
DEF_GEMM(f32, float)
DEF_GEMM(f64, double)
DEF_GEMM(i32, int32_t)
DEF_GEMM(i64, int64_t)

// dims = {N,C,H,W,Cout,kH,kW,padH,padW,sH,sW,dH,dW}
static ConvDims mk(const int64_t* p) {
  return ConvDims{p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10], p[11], p[12]};
}

#define DEF_CONV(SUF, T)                                                                               \
  EXPORT void oracle_im2col_##SUF(const T* in, const int64_t* dims, T* out) {                         \
    im2col<T>(in, mk(dims), out);                                                                      \
  }                                                                                                    \
  EXPORT void oracle_col2im_##SUF(const T* cols, const int64_t* dims, T* out) {                       \
    col2im<T>(cols, mk(dims), out);                                                                    \
  }                                                                                                    \
  EXPORT void oracle_conv2d_forward_##SUF(const T* in, const T* k, const T* bias, T* out,              \
                                          const int64_t* dims, int variant, int threads) {             \
    conv2d_forward<T>(in, k, bias, out, mk(dims), variant, threads);                                   \
  }                                                                                                    \
  EXPORT void oracle_conv2d_backward_##SUF(const T* in, const T* k, const T* gout, T* gin, T* gw,      \
                                           T* gb, const int64_t* dims, int variant, int threads) {     \
    conv2d_backward<T>(in, k, gout, gin, gw, gb, mk(dims), variant, threads);                          \
  }
DEF_CONV(f32, float)
DEF_CONV(f64, double)
DEF_CONV(i32, int32_t)
DEF_CONV(i64, int64_t)

EXPORT int oracle_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
