// Shared helpers for the B200 dense-contraction kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "../../include/am_b200.h"

namespace am {

// ---- status / last error (no exceptions or aborts cross the C ABI) -------------------------
void set_last_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);   // records + returns AM_ERR_CUDA

#define AM_CUDA_TRY(expr)                                  \
  do {                                                     \
    cudaError_t _e = (expr);                               \
    if (_e != cudaSuccess) return am::cuda_fail(_e, #expr); \
  } while (0)

// ---- per-device workspace cache (the shim may keep one, freed by am_shutdown) --
// slot ids
enum WsSlot : int { kWsSplitA = 0, kWsSplitB = 1, kWsConv = 2, kWsConvTab = 3, kWsMisc = 4, kWsConvW = 5, kWsNn = 6,
                    kWsStrideIn = 7, kWsStrideK = 8, kWsStrideGo = 9, kWsStrideOut = 10, kWsStrideGk = 11, kWsSkinny = 12, kWsGoutSplit = 13,
                    kWsNumSlots = 16 };
// returns device pointer valid until the next workspace() call for the same (device,slot) with a larger size
int workspace(int slot, size_t bytes, void** out);
void workspace_release_all();
int sm_count();

// ---- arithmetic that matches the reference's element semantics -----------------------------
// integers wrap mod 2^n: do everything in the unsigned type.
template <class T> struct UnsignedOf { using type = T; };
template <> struct UnsignedOf<int32_t> { using type = uint32_t; };
template <> struct UnsignedOf<int64_t> { using type = uint64_t; };

template <class T>
__device__ __forceinline__ T mac(T a, T b, T acc) {
  if constexpr (std::is_same<T, float>::value) return fmaf(a, b, acc);
  else if constexpr (std::is_same<T, double>::value) return fma(a, b, acc);
  else {
    using U = typename UnsignedOf<T>::type;
    return (T)((U)acc + (U)a * (U)b);
  }
}
// int64 multiply-accumulate mod 2^64 in exactly three IMAD-class instructions
// (IMAD.WIDE.U32 with 64-bit accumulate + two 32-bit IMADs into the high word); left to itself
// nvcc emits five (separate 64-bit add), measured on the first GEMM profile (profiles/r01_*).
template <>
__device__ __forceinline__ int64_t mac<int64_t>(int64_t a, int64_t b, int64_t acc) {
  const uint32_t alo = (uint32_t)a, ahi = (uint32_t)((uint64_t)a >> 32);
  const uint32_t blo = (uint32_t)b, bhi = (uint32_t)((uint64_t)b >> 32);
  uint64_t r;
  asm("{\n\t.reg .u32 lo, hi;\n\t"
      "mad.wide.u32 %0, %1, %2, %5;\n\t"
      "mov.b64 {lo, hi}, %0;\n\t"
      "mad.lo.u32 hi, %1, %4, hi;\n\t"
      "mad.lo.u32 hi, %3, %2, hi;\n\t"
      "mov.b64 %0, {lo, hi};\n\t}"
      : "=l"(r) : "r"(alo), "r"(blo), "r"(ahi), "r"(bhi), "l"((uint64_t)acc));
  return (int64_t)r;
}
// both operands known to fit in int32 (sign-extended): one IMAD.WIDE per multiply-accumulate, still exact mod 2^64
__device__ __forceinline__ int64_t mac_narrow_i64(int64_t a, int64_t b, int64_t acc) {
  int64_t r;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"((int32_t)a), "r"((int32_t)b), "l"(acc));
  return r;
}

template <class T>
__device__ __forceinline__ T mul_nocontract(T a, T b) {
  if constexpr (std::is_same<T, float>::value) return __fmul_rn(a, b);
  else if constexpr (std::is_same<T, double>::value) return __dmul_rn(a, b);
  else {
    using U = typename UnsignedOf<T>::type;
    return (T)((U)a * (U)b);
  }
}
template <class T>
__device__ __forceinline__ T add_nocontract(T a, T b) {
  if constexpr (std::is_same<T, float>::value) return __fadd_rn(a, b);
  else if constexpr (std::is_same<T, double>::value) return __dadd_rn(a, b);
  else {
    using U = typename UnsignedOf<T>::type;
    return (T)((U)a + (U)b);
  }
}

// Epilogue of gemm_ukernel_generic.nim:96-125 applied once to the full-K sum:
//   beta == 0 : C = (alpha == 1 ? AB : alpha*AB)          (never reads C)
//   else      : C = C*beta ; C += (alpha == 1 ? AB : alpha*AB)   (separately rounded)
template <class T>
__device__ __forceinline__ T epilogue_value(T alpha, T ab, T beta, T cold) {
  T v = (alpha == T(1)) ? ab : mul_nocontract<T>(alpha, ab);
  if (beta == T(0)) return v;
  T c = (beta == T(1)) ? cold : mul_nocontract<T>(cold, beta);
  return add_nocontract<T>(c, v);
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }
static inline int64_t iabs64(int64_t a) { return a < 0 ? -a : a; }
__device__ __forceinline__ int64_t iabs64_dev(int64_t a) { return a < 0 ? -a : a; }

}  // namespace am
