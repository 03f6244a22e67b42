/* am_b200.h — C ABI of libarraymancer_b200.so: the B200 (sm_100a) drop-in for Arraymancer's
 * dense-contraction hot path.  Plain pointers and sizes only; every entry point returns an
 * int status (0 = AM_OK) so the Nim side can wrap it in `check` exactly like it wraps cuBLAS
 * (reference: src/arraymancer/tensor/backend/cublas.nim:155-170).  No exceptions or aborts
 * cross this boundary; am_last_error() returns the message of the last failure on the
 * calling thread.  All device entry points are asynchronous on the given stream and never
 * free or retain user pointers.  All paths are relative to /root/reference/src/arraymancer/.
 * Concurrency: like the reference's single cudaStream0 / cublasHandle0 (backend/cuda_global_state.nim:21-39) the
 * library keeps ONE set of scratch buffers per device (packed operand panels, conv partials, ...), ordered by the
 * stream of the call that uses them: calls of the same family (GEMM, conv, NN ops) on one device must be issued on
 * one stream (or be ordered by events); different devices and different families are independent.
 * CUDA graphs: the device entries only enqueue kernels / copies on the given stream (no events, no other streams, no host
 * synchronisation), so after one eager call has sized the scratch buffers they can be stream-captured and replayed.
 *
 * The reference-side bindings (Nim {.importc, cdecl, dynlib.}) are shown in INTEGRATION.md.
 */
#ifndef AM_B200_H
#define AM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define AM_API
#else
#define AM_API __attribute__((visibility("default")))
#endif

typedef void* am_stream_t; /* a cudaStream_t (NULL = legacy default stream) */

enum {
  AM_OK = 0,
  AM_ERR_INVALID = 1,     /* bad argument (negative dim, null pointer with non-empty shape, ...) */
  AM_ERR_CUDA = 2,        /* a CUDA runtime/driver call failed; see am_last_error() */
  AM_ERR_UNSUPPORTED = 3, /* not built for this device (needs compute capability 10.x) */
  AM_ERR_NONCONTIGUOUS = 4 /* cuBLAS-shaped entry given a layout it cannot express */
};

/* ---- introspection ------------------------------------------------------------------- */
AM_API const char* am_version(void);
AM_API const char* am_last_error(void);
AM_API int am_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Frees the per-device workspace caches (split/packed operand panels, conv tables).
 * Reference scratch is per call (laser/.../gemm_tiling.nim:266-270,326); here it is cached. */
AM_API int am_shutdown(void);
/* Tuning knobs, set explicitly by the caller (process-wide atomics; the library reads NO environment variables):
 *   "tc_flush_kb" (2)  k-blocks of 32 per tensor-core accumulation chain of the 3xTF32 GEMM and conv wgrad kernels
 *   "tc_group" (8)     rasterisation group of the persistent GEMM (< 0: groups of N tiles)
 *   "tc_sync" (1)      per-wave grid barrier of the persistent GEMM
 *   "pack_scalar" (0)  force the scalar split/pack kernel
 *   "host_rowchunks" (0)  am_host_gemm_strided_f32 uses row chunks only (no K pipeline)
 *   "convtc_groups" (0 = kernel default), "convtc_debug" (0), "convtc_dgrad_gather" (0): tcgen05 conv kernels
 *   "simt_vec_load" (1)  128-bit global accesses along a unit-stride operand dimension in the SIMT GEMM
 *   "dmma_tma" (1)     TMA-fed 16-warp DMMA kernel for unit-stride float64 operands (0: register-staged 8-warp kernel)
 *   "convtc_hi_resident" (1)  tcgen05 conv forward keeps the hi weight planes resident in shared memory when they fit
 *   "convtc_flush_kb" (4)  k-blocks of 32 per accumulation chain of the tcgen05 conv forward (cv2: 7e-7 rel. Frobenius at 4,
 *                      4e-7 at 2 and 12 % slower); "convtc_debug" bits 1-2 switch pieces of that kernel off (experiments only)
 * Unknown names return AM_ERR_INVALID. */
AM_API int am_set_tuning(const char* name, int value);
AM_API int am_get_tuning(const char* name, int* value);

/* ---- GEMM: mirror of laser gemm_strided ------------------------------------------------
 * Replaces: laser/primitives/matrix_multiplication/gemm.nim:192-201 (`gemm_strided[T]`, raw
 * `ptr T` + (rowStride, colStride) in ELEMENTS, any sign / zero) and its unsigned/`int`
 * overload :275-307 (callers bit-cast uint32/uint64/int to i32/i64, as the reference does).
 * C <- alpha*A*B + beta*C with A[M,K], B[K,N], C[M,N]; element (r,c) of X at X[r*rs + c*cs].
 * beta == 0 never reads C (uninitialised / NaN safe, gemm_ukernel_generic.nim:53-60,103-111).
 * K == 0 leaves C untouched even if beta != 1, like the reference (gemm.nim:203 TODO).
 * Integers wrap mod 2^n.  Pointers are DEVICE pointers; `stream` is added first.
 * Callers today: tensor/private/p_operator_blas_l2l3.nim:131 (ints), and after integration
 * tensor/operators_blas_l2l3.nim:58-71 (floats), nn_primitives/fallback/conv.nim:103,136,140. */
AM_API int am_gemm_strided_f32(am_stream_t stream, int64_t M, int64_t N, int64_t K, float alpha,
                               const float* A, int64_t rowStrideA, int64_t colStrideA,
                               const float* B, int64_t rowStrideB, int64_t colStrideB, float beta,
                               float* C, int64_t rowStrideC, int64_t colStrideC);
AM_API int am_gemm_strided_f64(am_stream_t stream, int64_t M, int64_t N, int64_t K, double alpha,
                               const double* A, int64_t rowStrideA, int64_t colStrideA,
                               const double* B, int64_t rowStrideB, int64_t colStrideB, double beta,
                               double* C, int64_t rowStrideC, int64_t colStrideC);
AM_API int am_gemm_strided_i32(am_stream_t stream, int64_t M, int64_t N, int64_t K, int32_t alpha,
                               const int32_t* A, int64_t rowStrideA, int64_t colStrideA,
                               const int32_t* B, int64_t rowStrideB, int64_t colStrideB, int32_t beta,
                               int32_t* C, int64_t rowStrideC, int64_t colStrideC);
AM_API int am_gemm_strided_i64(am_stream_t stream, int64_t M, int64_t N, int64_t K, int64_t alpha,
                               const int64_t* A, int64_t rowStrideA, int64_t colStrideA,
                               const int64_t* B, int64_t rowStrideB, int64_t colStrideB, int64_t beta,
                               int64_t* C, int64_t rowStrideC, int64_t colStrideC);

/* Float32 path selector for am_gemm_strided_f32 (process-wide, default AM_F32_AUTO):
 * AUTO   = tcgen05 3xTF32 (split/packed panels -> TMA -> UMMA, fp32 accumulate in TMEM) for
 *          shapes that fill tensor tiles, exact-FFMA SIMT kernel for small / skinny shapes;
 * SIMT   = always the FFMA kernel;  TC = always tcgen05 3xTF32 (any shape, padded);
 * TC_1CTA = tcgen05 with cta_group::1 tiles (bring-up / comparison). */
enum { AM_F32_AUTO = 0, AM_F32_SIMT = 1, AM_F32_TC = 2, AM_F32_TC_1CTA = 3 };
AM_API int am_set_f32_path(int path);
AM_API int am_get_f32_path(void);

/* Float64 path selector for am_gemm_strided_f64 (process-wide, default AM_F64_AUTO):
 * AUTO = DMMA (mma.sync m8n8k4 f64, FP64 tensor pipe) once the output fills 128x128 tiles, DFMA
 * SIMT kernel for small shapes; SIMT / DMMA force one kernel. */
enum { AM_F64_AUTO = 0, AM_F64_SIMT = 1, AM_F64_DMMA = 2 };
AM_API int am_set_f64_path(int path);
AM_API int am_get_f64_path(void);

/* ---- pre-packed float32 operands ---------------------------------------------------------
 * Device-side counterpart of laser's pre-packed GEMM API
 * (laser/primitives/matrix_multiplication/gemm_prepacked.nim:276-293 `gemm_packed`, with
 * `gemm_prepackA/B` :178-270): split an operand once into its two K-major tf32 planes and
 * reuse it over many products (e.g. B across the row chunks of a sharded GEMM, or a weight
 * matrix across batches).  am_pack_f32_a packs A[M,K] (rows = m), am_pack_f32_b packs B[K,N]
 * (rows = n); handles own their device memory until am_packed_free_f32.  am_repack_f32_*
 * refreshes a handle in place from new data of the same shape.
 * am_gemm_packed_f32: C[M,N] <- alpha*A*B + beta*C on the tcgen05 3xTF32 path. */
typedef struct am_packed_f32 am_packed_f32;
AM_API int am_pack_f32_a(am_stream_t stream, int64_t M, int64_t K, const float* A, int64_t rowStrideA,
                         int64_t colStrideA, am_packed_f32** out);
AM_API int am_pack_f32_b(am_stream_t stream, int64_t K, int64_t N, const float* B, int64_t rowStrideB,
                         int64_t colStrideB, am_packed_f32** out);
AM_API int am_repack_f32_a(am_stream_t stream, am_packed_f32* h, const float* A, int64_t rowStrideA,
                           int64_t colStrideA);
AM_API int am_repack_f32_b(am_stream_t stream, am_packed_f32* h, const float* B, int64_t rowStrideB,
                           int64_t colStrideB);
AM_API int am_gemm_packed_f32(am_stream_t stream, float alpha, const am_packed_f32* A,
                              const am_packed_f32* B, float beta, float* C, int64_t rowStrideC,
                              int64_t colStrideC);
/* Row-sharded GEMM fused with the all-gather of its result ("rank r owns rows of A and of C ...
 * ncclAllGather of C"): C <- alpha*A*B is stored by the GEMM epilogue to the same offsets of EVERY GPU's copy
 * of C.  peerC[g] = device address, valid in the calling process (peer-mapped / symmetric memory over NVLink),
 * of element (0,0) of the output block in GPU g's buffer; peerC[self_index] is the caller's own copy (written in
 * the tile epilogue; the copies to the peers are re-read from it and posted between the accumulation chains of the
 * following tile, so the NVLink traffic is spread over the mainloop instead of bursting).
 * beta is 0 by construction (remote copies are write-only).  1 <= npeers <= 8.  The caller synchronises the
 * GPUs (stream order + a cross-GPU barrier) before any of them reads C. */
AM_API int am_gemm_packed_f32_bcast(am_stream_t stream, float alpha, const am_packed_f32* A,
                                    const am_packed_f32* B, int npeers, float* const* peerC,
                                    int self_index, int64_t rowStrideC, int64_t colStrideC);
AM_API int am_packed_free_f32(am_packed_f32* h);
/* Packed operands in CALLER-OWNED device memory (the handle is only a descriptor; am_packed_free_f32 drops it and
 * leaves the memory alone): am_packed_floats_f32(R, K) = floats a packed operand of R panel rows (M for an A, N for a
 * B) and depth K occupies; am_pack_f32_{a,b}_into pack into `planes`; am_packed_wrap_f32 describes planes that already
 * hold a packed operand of that shape — e.g. a K slice of B split/packed by another GPU and received over NVLink
 * (the multi-GPU host-buffer GEMM exchanges packed K slices instead of uploading all of B on every GPU). */
AM_API int64_t am_packed_floats_f32(int64_t R, int64_t K);
AM_API int am_pack_f32_a_into(am_stream_t stream, int64_t M, int64_t K, const float* A, int64_t rowStrideA,
                              int64_t colStrideA, float* planes, am_packed_f32** out);
AM_API int am_pack_f32_b_into(am_stream_t stream, int64_t K, int64_t N, const float* B, int64_t rowStrideB,
                              int64_t colStrideB, float* planes, am_packed_f32** out);
AM_API int am_packed_wrap_f32(int64_t R, int64_t K, float* planes, am_packed_f32** out);

/* ---- batched GEMM ----------------------------------------------------------------------------
 * Replaces: tensor/backend/cublas.nim:172-208 `cublas_gemmStridedBatched` (declared, no caller yet) and serves the
 * reference's "TODO: batch matmul" over the images of a conv (nn_primitives/fallback/conv.nim:99).
 * For b in [0, batch): C_b <- alpha*A_b*B_b + beta*C_b with X_b = X + b*batchStrideX (elements; 0 = the operand is
 * shared by all products), every product with the semantics of am_gemm_strided_*.  Small products share one launch
 * (blockIdx.z = b); products that fill the chip on their own take the single-product kernels one after the other. */
#define AM_DECL_BATCHED(SUF, T)                                                                                  \
  AM_API int am_gemm_strided_batched_##SUF(am_stream_t stream, int64_t batch, int64_t M, int64_t N, int64_t K, T alpha, \
                                           const T* A, int64_t rowStrideA, int64_t colStrideA, int64_t batchStrideA, \
                                           const T* B, int64_t rowStrideB, int64_t colStrideB, int64_t batchStrideB, \
                                           T beta, T* C, int64_t rowStrideC, int64_t colStrideC, int64_t batchStrideC);
AM_DECL_BATCHED(f32, float)
AM_DECL_BATCHED(f64, double)
AM_DECL_BATCHED(i32, int32_t)
AM_DECL_BATCHED(i64, int64_t)
#undef AM_DECL_BATCHED

/* ---- cuBLAS-shaped adapter -------------------------------------------------------------
 * Replaces: tensor/backend/cublas.nim:142-170 `cublas_gemm[T]` (column-major, op N/T), whose
 * only caller is tensor/operators_blas_l2l3_cuda.nim:68-72 (`cudaMM_C_eq_aAB_p_bC`).
 * transa/transb: 0 = CUBLAS_OP_N, 1 = CUBLAS_OP_T.  Thin adapter over am_gemm_strided_*:
 * op N -> (rs=1, cs=ld); op T -> (rs=ld, cs=1); C is always (1, ldc). */
AM_API int am_cublas_gemm_f32(am_stream_t stream, int transa, int transb, int64_t m, int64_t n,
                              int64_t k, float alpha, const float* A, int64_t lda, const float* B,
                              int64_t ldb, float beta, float* C, int64_t ldc);
AM_API int am_cublas_gemm_f64(am_stream_t stream, int transa, int transb, int64_t m, int64_t n,
                              int64_t k, double alpha, const double* A, int64_t lda,
                              const double* B, int64_t ldb, double beta, double* C, int64_t ldc);

/* ---- conv2d: fused implicit-GEMM (the im2col buffer is never materialised) ---------------
 * Replaces: nn_primitives/fallback/conv.nim:81-106 `im2colgemm_conv2d` and :108-140
 * `im2colgemm_conv2d_gradient` (+ nnp_convolution.nim:91-94 grad_bias), and on the GPU
 * boundary nn_primitives/nnp_conv2d_cudnn.nim:20-72 `conv2d` / :74-204 `conv2d_backward`.
 * Layout: input [N,C,H,W], kernel [Cout,C,kH,kW], output / grad_output [N,Cout,Ho,Wo], all
 * C-contiguous DEVICE buffers pre-allocated by the caller; bias [Cout] or NULL (rank-0 bias).
 * Ho = (H + 2*padH - (dH*(kH-1)+1)) / sH + 1 (conv.nim:90-91; dilation as in the cuDNN
 * signature).  Cross-correlation (no kernel flip). */
typedef struct am_conv2d_desc {
  int64_t N, C, H, W;
  int64_t Cout, kH, kW;
  int64_t padH, padW, strideH, strideW, dilH, dilW;
} am_conv2d_desc;

AM_API int am_conv2d_out_dims(const am_conv2d_desc* d, int64_t* Ho, int64_t* Wo);
/* Conv kernel selector (process-wide, default AM_CONV_AUTO):
 * AUTO   = float32: per call the family that measured fastest for the shape — tcgen05 kernels (forward when
 *          Cout >= 32 and C*kH*kW >= 128; GEMM + col2im data gradient for single-tile images; weight gradient
 *          whenever it fits: Cout <= 64), else the direct SIMT kernels, else the generic gather kernels;
 *          other dtypes: generic gather kernels (bit-exact integers);
 * GATHER = generic implicit-GEMM gather kernels everywhere (fallback / comparison);
 * DIRECT = float32 direct SIMT kernels where they apply (stride-1, dilation-1, kW in {1,3,5,7}), no tensor cores;
 * TC     = float32 tcgen05 kernels wherever they fit (Cout <= 64, ...), then DIRECT, then GATHER.
 * Every family is deterministic and is run over the whole parity matrix (tests/test_gpu_conv.py). */
enum { AM_CONV_AUTO = 0, AM_CONV_GATHER = 1, AM_CONV_DIRECT = 2, AM_CONV_TC = 3 };
AM_API int am_set_conv_path(int path);

/* Epilogue fusion on the conv boundary: the reference adds the bias in a second pass
 * (fallback/conv.nim:105-106, nnp_conv2d_cudnn.nim:72) and applies relu in a third (nnp_activation.nim:35-36);
 * am_conv2d_forward_* fuses the bias, am_conv2d_forward_act_* also the activation: y = max(0, conv + bias) with
 * the reference's relu semantics (value <= 0 -> 0, NaN stays NaN).  relu_backward on the stored output is
 * equivalent to relu_backward on the pre-activation (output <= 0 exactly where pre-activation <= 0). */
enum { AM_ACT_NONE = 0, AM_ACT_RELU = 1 };
#define AM_DECL_CONV(SUF, T)                                                                     \
  AM_API int am_conv2d_forward_##SUF(am_stream_t stream, const am_conv2d_desc* d, const T* input, \
                                     const T* kernel, const T* bias, T* output);                 \
  AM_API int am_conv2d_forward_act_##SUF(am_stream_t stream, const am_conv2d_desc* d, const T* input, \
                                         const T* kernel, const T* bias, T* output, int activation); \
  /* grad_input / grad_kernel / grad_bias may each be NULL to skip that gradient */             \
  AM_API int am_conv2d_backward_##SUF(am_stream_t stream, const am_conv2d_desc* d,               \
                                      const T* input, const T* kernel, const T* grad_output,     \
                                      T* grad_input, T* grad_kernel, T* grad_bias);
AM_DECL_CONV(f32, float)
AM_DECL_CONV(f64, double)
AM_DECL_CONV(i32, int32_t)
AM_DECL_CONV(i64, int64_t)
#undef AM_DECL_CONV

/* conv2d on STRIDED tensors: every tensor comes with its 4 element strides, the way the reference builds its
 * descriptors from `t.strides[0..3]` (nn_primitives/backend/cudnn.nim:59-75), so views (transposed, sliced, Fortran-
 * ordered grad_output ...) cross the boundary as they are — the Nim `conv2d_backward` no longer needs the
 * `asContiguous` that is `{.error.}` on CUDA (nnp_conv2d_cudnn.nim:99-101).  A NULL strides pointer
 * means dense NCHW; bias / grad_bias take one element stride (1 = dense [Cout]).  Dense tensors go straight to the
 * kernels of am_conv2d_*; a non-dense tensor costs one extra pass over it (gathered into / scattered from a dense
 * workspace copy by a coalescing copy kernel).  Same NULL-gradient convention as am_conv2d_backward_*. */
#define AM_DECL_CONV_STRIDED(SUF, T)                                                                             \
  AM_API int am_conv2d_forward_strided_##SUF(am_stream_t stream, const am_conv2d_desc* d, const T* input,         \
                                             const int64_t* input_strides, const T* kernel,                       \
                                             const int64_t* kernel_strides, const T* bias, int64_t bias_stride,   \
                                             T* output, const int64_t* output_strides, int activation);           \
  AM_API int am_conv2d_backward_strided_##SUF(am_stream_t stream, const am_conv2d_desc* d, const T* input,        \
                                              const int64_t* input_strides, const T* kernel,                      \
                                              const int64_t* kernel_strides, const T* grad_output,                \
                                              const int64_t* grad_output_strides, T* grad_input,                  \
                                              const int64_t* grad_input_strides, T* grad_kernel,                  \
                                              const int64_t* grad_kernel_strides, T* grad_bias,                   \
                                              int64_t grad_bias_stride);
AM_DECL_CONV_STRIDED(f32, float)
AM_DECL_CONV_STRIDED(f64, double)
AM_DECL_CONV_STRIDED(i32, int32_t)
AM_DECL_CONV_STRIDED(i64, int64_t)
#undef AM_DECL_CONV_STRIDED

/* ---- host-buffer entry points (the reference-facing "plugin" call with HOST memory) ------
 * What `a.cuda * b.cuda` followed by `.cpu` does in the reference
 * (tensor/init_cuda.nim:23-59 + operators_blas_l2l3_cuda.nim:74-87): device buffers are
 * allocated, operands copied H2D on the call's stream, the GEMM enqueued, the result copied
 * D2H; returns after the result is in `C` (synchronises the stream).  Strides in elements,
 * any layout the device entry accepts.  Used by bench.py's `e2e` measurement. */
AM_API int am_host_gemm_strided_f32(int64_t M, int64_t N, int64_t K, float alpha, const float* A,
                                    int64_t rsA, int64_t csA, const float* B, int64_t rsB,
                                    int64_t csB, float beta, float* C, int64_t rsC, int64_t csC);
AM_API int am_host_gemm_strided_f64(int64_t M, int64_t N, int64_t K, double alpha, const double* A,
                                    int64_t rsA, int64_t csA, const double* B, int64_t rsB,
                                    int64_t csB, double beta, double* C, int64_t rsC, int64_t csC);
AM_API int am_host_gemm_strided_i32(int64_t M, int64_t N, int64_t K, int32_t alpha, const int32_t* A,
                                    int64_t rsA, int64_t csA, const int32_t* B, int64_t rsB,
                                    int64_t csB, int32_t beta, int32_t* C, int64_t rsC, int64_t csC);
AM_API int am_host_gemm_strided_i64(int64_t M, int64_t N, int64_t K, int64_t alpha, const int64_t* A,
                                    int64_t rsA, int64_t csA, const int64_t* B, int64_t rsB,
                                    int64_t csB, int64_t beta, int64_t* C, int64_t rsC, int64_t csC);

/* Pitched copy on `stream` (cudaMemcpy2DAsync): a column block / K slice of a row-major HOST matrix to or from device
 * memory without staging.  Pitches and width in BYTES; kind: 1 = host->device, 2 = device->host, 3 = device->device
 * (UVA: also peer-mapped memory of another GPU, moved by the copy engines over NVLink).  Host memory should be pinned. */
AM_API int am_memcpy2d_async(am_stream_t stream, void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch,
                             int64_t width, int64_t rows, int kind);

/* ---- multi-GPU (one host process drives the GPUs of a box) ------------------------------------------------
 * The reference has no multi-device code and never calls cudaSetDevice; these entries manage the devices
 * themselves and restore the caller's current device before returning (last row).  am_mg_init enables peer
 * access between the listed devices (NULL = 0..ndev-1; 1 <= ndev <= 8) and creates the context's streams.
 * Row-sharded GEMM: GPU g owns the contiguous rows [row0_g, row0_g + rows_g) of A and of C as reported by
 * am_mg_rows(ctx, M, g, ...); A_local[g] points at ITS rows (rows_g x K, row pitch ldA) in GPU g's memory, B[g] at GPU
 * g's replica of B (K x N, pitch ldB), C[g] at GPU g's buffer for the WHOLE result (M x N, pitch ldC).  After the call
 * (asynchronous: am_mg_synchronize, or enqueue more work on am_mg_stream(ctx, g)) every C[g] holds alpha*A*B.
 *   f32: the tcgen05 GEMM of every GPU stores its tiles into every GPU's copy of C from its epilogue, over NVLink
 *        (peer-mapped pointers): product and all-gather are one kernel;
 *   f64 / i32 / i64: row chunks, each pushed to the peers by the copy engines while the next chunk computes.
 * No K split: integers stay bit-exact, floats are exactly the single-GPU kernels' results.
 * am_mg_host_gemm_f32: A, B, C in HOST memory (row-major, pinned for speed); GPU g uploads its rows of A and only a
 * 1/ndev share of B, the shares are exchanged between the GPUs; synchronous. */
typedef struct am_mg_ctx am_mg_ctx;
AM_API int am_mg_init(int ndev, const int* devices, am_mg_ctx** out);
AM_API int am_mg_destroy(am_mg_ctx* ctx);
AM_API int am_mg_device_count(const am_mg_ctx* ctx);
AM_API void* am_mg_stream(const am_mg_ctx* ctx, int g);
AM_API int am_mg_rows(const am_mg_ctx* ctx, int64_t M, int g, int64_t* row0, int64_t* rows);
AM_API int am_mg_synchronize(am_mg_ctx* ctx);
#define AM_DECL_MG(SUF, T)                                                                                        \
  AM_API int am_mg_gemm_rowsharded_##SUF(am_mg_ctx* ctx, int64_t M, int64_t N, int64_t K, T alpha,                 \
                                         const T* const* A_local, int64_t ldA, const T* const* B, int64_t ldB,     \
                                         T* const* C, int64_t ldC);
AM_DECL_MG(f32, float)
AM_DECL_MG(f64, double)
AM_DECL_MG(i32, int32_t)
AM_DECL_MG(i64, int64_t)
#undef AM_DECL_MG
AM_API int am_mg_host_gemm_f32(am_mg_ctx* ctx, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t ldA,
                               const float* B, int64_t ldB, float* C, int64_t ldC);

/* ---- LeNet companions: the HBM-bound operators between the contractions ------------
 * so that a forward + backward step of the reference's ex02_mnist network stays on the device.  Dense NCHW /
 * row-major device buffers, outputs pre-allocated by the caller, asynchronous on `stream`, deterministic.
 *   relu / relu_backward             nn_primitives/nnp_activation.nim:35-36, 65-70  (max(0,x); cached <= 0 ? 0 : g)
 *   maxpool2d / maxpool2d_backward   nn_primitives/nnp_maxpooling.nim:19-83  (max_indices = flat input index, int64;
 *                                    first maximum wins; backward ASSIGNS grad_out[i] to grad_in[max_indices[i]],
 *                                    the last i wins where windows overlap: pass windows_overlap = 1 unless
 *                                    stride >= kernel in both dimensions)
 *   maxpool2d_backward_relu          = relu_backward(maxpool2d_backward(grad_out), relu_cached), bit for bit, in ONE pass:
 *                                    the mask of nnp_activation.nim:65-70 is applied to the one element per window that
 *                                    receives a gradient (conv -> relu -> maxpool blocks of ex02_mnist.nim: the
 *                                    element-wise relu_backward pass over the conv output disappears)
 *   linear / linear_backward         nn_primitives/nnp_linear.nim:20-66  (y = x*W^T + b; gI = gO*W, gW = gO^T*x,
 *                                    gB = sum(gO, axis 0); bias / any gradient pointer may be NULL)
 *   sparse_softmax_cross_entropy     nn_primitives/nnp_softmax_cross_entropy.nim:100-178 (mean over the batch of
 *   (+ _backward)                    logsumexp(x_i) - x_i[label_i], streaming max / sum-exp per row; the scalar loss
 *                                    is written to device memory) and :219-252 (grad*(softmax - onehot)/batch). */
#define AM_DECL_NN(SUF, T)                                                                                 \
  AM_API int am_relu_forward_##SUF(am_stream_t stream, int64_t n, const T* x, T* y);                         \
  AM_API int am_relu_backward_##SUF(am_stream_t stream, int64_t n, const T* gradient, const T* cached, T* out); \
  AM_API int am_maxpool2d_forward_##SUF(am_stream_t stream, int64_t N, int64_t C, int64_t H, int64_t W, int64_t kH, \
                                        int64_t kW, int64_t padH, int64_t padW, int64_t strideH, int64_t strideW,  \
                                        const T* input, T* maxpooled, int64_t* max_indices);                 \
  AM_API int am_maxpool2d_backward_##SUF(am_stream_t stream, int64_t n_input, int64_t n_output,              \
                                         const int64_t* max_indices, const T* grad_output, T* grad_input,    \
                                         int windows_overlap);                                               \
  AM_API int am_maxpool2d_backward_relu_##SUF(am_stream_t stream, int64_t n_input, int64_t n_output,         \
                                              const int64_t* max_indices, const T* grad_output,              \
                                              const T* relu_cached, T* grad_input, int windows_overlap);     \
  AM_API int am_linear_forward_##SUF(am_stream_t stream, int64_t batch, int64_t in_features, int64_t out_features, \
                                     const T* input, const T* weight, const T* bias, T* output);             \
  AM_API int am_linear_backward_##SUF(am_stream_t stream, int64_t batch, int64_t in_features, int64_t out_features, \
                                      const T* input, const T* weight, const T* grad_output, T* grad_input,  \
                                      T* grad_weight, T* grad_bias);                                         \
  AM_API int am_sparse_softmax_cross_entropy_##SUF(am_stream_t stream, int64_t batch, int64_t features,      \
                                                   const T* input, int64_t rowStride, int64_t colStride,     \
                                                   const int64_t* labels, T* loss_device);                   \
  AM_API int am_sparse_softmax_cross_entropy_backward_##SUF(am_stream_t stream, int64_t batch, int64_t features, \
                                                            T gradient, const T* cached_input, int64_t rowStride, \
                                                            int64_t colStride, const int64_t* labels, T* grad_input);
AM_DECL_NN(f32, float)
AM_DECL_NN(f64, double)
#undef AM_DECL_NN

/* ---- measurement helpers (bench.py / profiles) -------------------------------------------
 * Number of kernels this library has launched on the calling process since load. */
AM_API int64_t am_kernel_launch_count(void);
/* On-device micro-benchmarks that define the per-dtype roofline denominators:
 * which: 0 = FFMA f32, 1 = DFMA f64, 2 = IMAD i32, 3 = i64 multiply-add sequence,
 *        4 = DMMA m8n8k4 f64 (mma.sync), 5 = tcgen05 kind::tf32 cta_group::1 128x256x8,
 *        6 = tcgen05 kind::tf32 cta_group::2 256x256x8, 7 = int64 accumulate of int32-range operands
 *        (one IMAD.WIDE per multiply-accumulate); 8..12 = tcgen05 kind::tf32 cta_group::1 at the small N the
 *        implicit-GEMM conv issues (compile-time shapes, 16 unrolled MMAs per commit): 8 = 128x64x8 into one
 *        accumulator, 9 = alternating two accumulators,
 *        10 = 128x64x8 with A read from tensor memory, 11 = same with two accumulators, 12 = 128x128x8.
 * Writes achieved 1e12 op/s (2 ops per multiply-add) to *tops. */
AM_API int am_microbench(int which, double* tops);

#ifdef __cplusplus
}
#endif
#endif /* AM_B200_H */
