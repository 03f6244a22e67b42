// Strided SIMT GEMM, int64_t instantiation (one translation unit per element type: the staging-mode x tile x batched
// matrix of contract_simt_kernel compiles in parallel).
#include "gemm_simt_impl.cuh"

namespace am {

std::atomic<int64_t> g_launch_count{0};

// Pre-pass of the int64 GEMM: is every element of A and of B representable in int32?  (flag := 1 if not.)
// O(MK + KN) reads; lets the mainloop use one IMAD.WIDE per multiply-accumulate (bit-identical result).
__global__ void i64_range_kernel(const int64_t* __restrict__ A, int64_t a_mn, int64_t a_k, int64_t M,
                                 const int64_t* __restrict__ B, int64_t b_mn, int64_t b_k, int64_t N, int64_t K,
                                 int* __restrict__ wide_flag) {
  const bool isB = blockIdx.y == 1;
  const int64_t* X = isB ? B : A;
  const int64_t mn_stride = isB ? b_mn : a_mn, k_stride = isB ? b_k : a_k, MN = isB ? N : M;
  const bool k_fast = iabs64_dev(k_stride) <= iabs64_dev(mn_stride);
  const int64_t inner = k_fast ? K : MN, total = MN * K;
  bool wide = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = i / inner, in = i - o * inner;
    const int64_t v = k_fast ? X[o * mn_stride + in * k_stride] : X[in * mn_stride + o * k_stride];
    wide |= (v != (int64_t)(int32_t)v);
  }
  if (__any_sync(0xffffffffu, wide) && (threadIdx.x & 31) == 0) atomicOr(wide_flag, 1);
}


AM_INST_SIMT(int64_t)

}  // namespace am
