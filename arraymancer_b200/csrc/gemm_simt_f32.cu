// Strided SIMT GEMM, float instantiation (one translation unit per element type: the staging-mode x tile x batched
// matrix of contract_simt_kernel compiles in parallel).
#include "gemm_simt_impl.cuh"

namespace am {

AM_INST_SIMT(float)

}  // namespace am
