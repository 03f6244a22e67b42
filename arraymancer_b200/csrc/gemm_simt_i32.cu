// Strided SIMT GEMM, int32_t instantiation (one translation unit per element type: the staging-mode x tile x batched
// matrix of contract_simt_kernel compiles in parallel).
#include "gemm_simt_impl.cuh"

namespace am {

AM_INST_SIMT(int32_t)

}  // namespace am
