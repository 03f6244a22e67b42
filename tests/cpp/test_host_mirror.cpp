// C++ twin of the reference's own CUDA tests, driven through the C++ host mirror (arraymancer_b200.hpp):
//   tests/tensor/test_operators_blas_cuda.nim:20-98        `a.cuda * b.cuda` known answers, transposes
//   tests/nn_primitives/test_nnp_convolution.nim:21-134    conv2d known answers (int exact, float32)
//   laser/primitives/matrix_multiplication/gemm.nim:394-418 integer self-test with negatives
// Prints "OK <n> checks" and exits 0, or the first failure and exits 1.  Needs a GPU.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../arraymancer_b200/host/arraymancer_b200.hpp"

using namespace arraymancer_b200;
static int checks = 0;
#define EXPECT(cond)                                                    \
  do {                                                                  \
    checks++;                                                           \
    if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } \
  } while (0)

template <class T>
static void gemm_known_answers() {
  // test_operators_blas_cuda.nim:22-33 / gemm.nim:369-392
  auto a = cuda<T>({1, 2, 3, 4, 5, 6}, {2, 3});
  auto b = cuda<T>({7, 8, 9, 10, 11, 12}, {3, 2});
  EXPECT((cpu(a * b) == std::vector<T>{58, 64, 139, 154}));
  // gemm.nim:394-418 — (M x K) * (K x N) with M < N, negatives
  auto u = cuda<T>({-2, -3, -1, 3, 0, 4}, {2, 3});
  auto v = cuda<T>({1, 5, 2, -1, -3, 0, 3, 4, 6, -2, 7, -4}, {3, 4});
  EXPECT((cpu(u * v) == std::vector<T>{1, -8, -20, -6, 27, 7, 34, -19}));
  // transposed views (test_operators_blas.nim:133-155)
  auto at = cuda<T>({1, 4, 2, 5, 3, 6}, {3, 2});
  auto bt = cuda<T>({7, 9, 11, 8, 10, 12}, {2, 3});
  EXPECT((cpu(at.transpose() * b) == std::vector<T>{58, 64, 139, 154}));
  EXPECT((cpu(a * bt.transpose()) == std::vector<T>{58, 64, 139, 154}));
  EXPECT((cpu(at.transpose() * bt.transpose()) == std::vector<T>{58, 64, 139, 154}));
  // alpha / beta through gemm(): C = 2*A*B + 3*C
  auto c = cuda<T>({1, 1, 1, 1}, {2, 2});
  gemm<T>(T(2), a, b, T(3), c);
  EXPECT((cpu(c) == std::vector<T>{119, 131, 281, 311}));
  // negative-step column slice of a column-major matrix (test_operators_blas.nim:157-193 pattern): b[:, ::-1]
  auto brev = b.slice(1, 1, 2, -1);
  EXPECT((cpu(a * brev) == std::vector<T>{64, 58, 154, 139}));
  // error conventions
  bool threw = false;
  try { (void)(a * a); } catch (const std::out_of_range&) { threw = true; }
  EXPECT(threw);
}

template <class T>
static CudaTensor<T> nchw(const std::vector<T>& data, const std::vector<int64_t>& shape) {
  CudaTensor<T> t = CudaTensor<T>::make(shape, /*colMajor=*/false);
  cudaCheck(cudaMemcpy(t.storage->data, data.data(), data.size() * sizeof(T), cudaMemcpyHostToDevice));
  return t;
}

template <class T>
static void conv_known_answers() {
  // test_nnp_convolution.nim:21-46
  auto x = nchw<T>({1, 2, 0, 0, 5, 3, 0, 4, 0, 0, 0, 7, 9, 3, 0, 0}, {1, 1, 4, 4});
  auto k = nchw<T>({1, 1, 1, 1, 1, 0, 1, 0, 0}, {1, 1, 3, 3});
  auto b = nchw<T>({0}, {1, 1, 1});
  auto y = conv2d<T>(x, k, b, {1, 1});
  std::vector<T> got((size_t)y.size());
  cudaCheck(cudaMemcpy(got.data(), y.storage->data, got.size() * sizeof(T), cudaMemcpyDeviceToHost));
  EXPECT((got == std::vector<T>{1, 8, 5, 0, 8, 11, 5, 4, 8, 17, 10, 11, 9, 12, 10, 7}));
  // test_nnp_convolution.nim:54-134 — 3 input channels, 2 filters, pad 1, stride 2, bias [1, 0]
  auto x2 = nchw<T>({2, 2, 0, 2, 1, 0, 1, 1, 0, 2, 1, 2, 1, 2, 1, 2, 2, 0, 0, 2, 2, 1, 1, 1, 2,
                     2, 0, 1, 1, 1, 2, 2, 0, 0, 2, 2, 2, 1, 0, 0, 1, 1, 2, 2, 0, 2, 1, 1, 1, 0,
                     0, 1, 2, 2, 0, 1, 1, 1, 1, 0, 2, 1, 2, 2, 0, 0, 2, 2, 2, 1, 0, 0, 2, 2, 1}, {1, 3, 5, 5});
  auto k2 = nchw<T>({-1, -1, -1, 1, 0, 1, 0, -1, 0, 1, 0, -1, 1, -1, 1, 0, 1, 0, 0, 0, 1, -1, -1, -1, -1, 0, 0,
                     0, 1, 0, 1, -1, -1, 1, 1, -1, -1, 0, 1, -1, -1, 1, 1, 1, 0, 0, 1, 1, -1, 1, -1, -1, -1, 0}, {2, 3, 3, 3});
  auto b2 = nchw<T>({1, 0}, {2, 1, 1});
  auto y2 = conv2d<T>(x2, k2, b2, {1, 1}, {2, 2});
  std::vector<T> got2((size_t)y2.size());
  cudaCheck(cudaMemcpy(got2.data(), y2.storage->data, got2.size() * sizeof(T), cudaMemcpyDeviceToHost));
  EXPECT((got2 == std::vector<T>{2, -2, 0, -3, 2, -5, -2, -1, 0, -7, 1, 0, 3, -3, 2, 1, 3, -2}));
  // backward runs and grad_bias with grad_output = ones equals N*Ho*Wo (test_nnp_convolution.nim:161)
  std::vector<T> ones((size_t)y2.size(), T(1));
  auto go = nchw<T>(ones, y2.shape);
  CudaTensor<T> gi, gk, gb;
  conv2d_backward<T>(x2, k2, b2, {1, 1}, {2, 2}, {1, 1}, go, gi, gk, gb);
  std::vector<T> gbh(2);
  cudaCheck(cudaMemcpy(gbh.data(), gb.storage->data, 2 * sizeof(T), cudaMemcpyDeviceToHost));
  EXPECT((gbh == std::vector<T>{9, 9}));
}

template <class T>
static std::vector<T> raw(const CudaTensor<T>& t) {
  std::vector<T> h((size_t)t.size());
  cudaCheck(cudaMemcpy(h.data(), t.get_offset_ptr(), h.size() * sizeof(T), cudaMemcpyDeviceToHost));
  return h;
}

template <class T>
static void nn_known_answers() {
  // tests/nn_primitives/test_nnp_maxpool.nim:21-32
  auto a = nchw<T>({1, 1, 2, 4, 5, 6, 7, 8, 3, 2, 1, 0, 1, 2, 3, 4}, {1, 1, 4, 4});
  auto mp = maxpool2d<T>(a, {2, 2}, {0, 0}, {2, 2});
  EXPECT((raw(mp.maxpooled) == std::vector<T>{6, 8, 3, 4}));
  EXPECT((raw(mp.max_indices) == std::vector<int64_t>{5, 7, 8, 15}));
  auto g = maxpool2d_backward<T>(a.shape, mp.max_indices, mp.maxpooled);
  EXPECT((raw(g) == std::vector<T>{0, 0, 0, 0, 0, 6, 0, 8, 3, 0, 0, 0, 0, 0, 0, 4}));
  // relu / relu_backward — nnp_activation.nim:35-36, 65-70
  auto x = nchw<T>({-1, 0, 2, -3}, {4});
  EXPECT((raw(relu(x)) == std::vector<T>{0, 0, 2, 0}));
  auto gr = nchw<T>({5, 6, 7, 8}, {4});
  EXPECT((raw(relu_backward(gr, x)) == std::vector<T>{0, 0, 7, 0}));
  // linear — nnp_linear.nim:20-66: y = x W^T + b on exact small integers
  auto xi = nchw<T>({1, 2, 3, 4, 5, 6}, {2, 3});
  auto w = nchw<T>({1, 0, -1, 2, 1, 0}, {2, 3});
  auto b = nchw<T>({10, 20}, {1, 2});
  EXPECT((raw(linear(xi, w, b)) == std::vector<T>{8, 24, 8, 33}));
  CudaTensor<T> gi, gw, gb;
  auto go = nchw<T>({1, 1, 2, 0}, {2, 2});
  linear_backward(xi, w, go, gi, gw, &gb);
  EXPECT((raw(gi) == std::vector<T>{3, 1, -1, 2, 0, -2}));
  EXPECT((raw(gw) == std::vector<T>{9, 12, 15, 1, 2, 3}));
  EXPECT((raw(gb) == std::vector<T>{3, 1}));
  // tests/nn_primitives/test_nnp_loss.nim:29-44: loss ~= 0.0709 (|a - b| <= 2e-5)
  auto pred = nchw<T>({T(-3.44), T(1.16), T(-0.81), T(3.91)}, {1, 4});
  auto lab = nchw<int64_t>({3}, {1});
  const T loss = sparse_softmax_cross_entropy(pred, lab);
  EXPECT(std::fabs((double)loss - 0.0709) <= 2e-5);
  auto gl = raw(sparse_softmax_cross_entropy_backward(T(1), pred, lab));
  double sum = 0;
  for (auto v : gl) sum += (double)v;
  EXPECT(std::fabs(sum) < 1e-6 && gl[3] < 0 && gl[0] > 0);      // softmax - onehot sums to zero
}

// Round-2 boundary entries straight through the C ABI: batched GEMM (cublas.nim:172-208), strided conv
// (backend/cudnn.nim:59-75 strides), single-process multi-GPU (am_mg_*; device 0 listed twice on a one-GPU box).
template <class T>
static T* dev_copy(const std::vector<T>& h) {
  T* d = nullptr;
  cudaCheck(cudaMalloc(&d, h.size() * sizeof(T)));
  cudaCheck(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}
template <class T>
static std::vector<T> host_copy(const T* d, size_t n) {
  std::vector<T> h(n);
  cudaCheck(cudaDeviceSynchronize());
  cudaCheck(cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost));
  return h;
}

static void boundary_round2() {
  // ---- batched: 3 products [2x3]*[3x2], B shared (batch stride 0); expected from gemm.nim:369-392 scaled per batch
  std::vector<int64_t> a;
  for (int b = 0; b < 3; b++) for (int v : {1, 2, 3, 4, 5, 6}) a.push_back((int64_t)v * (b + 1));
  int64_t* dA = dev_copy(a);
  int64_t* dB = dev_copy(std::vector<int64_t>{7, 8, 9, 10, 11, 12});
  int64_t* dC = dev_copy(std::vector<int64_t>(12, -1));
  amCheck(am_gemm_strided_batched_i64(nullptr, 3, 2, 2, 3, 1, dA, 3, 1, 6, dB, 2, 1, 0, 0, dC, 2, 1, 4));
  EXPECT((host_copy(dC, 12) == std::vector<int64_t>{58, 64, 139, 154, 116, 128, 278, 308, 174, 192, 417, 462}));
  // ---- strided conv: test_nnp_convolution.nim:21-46 with the input stored TRANSPOSED (W-major) and passed by strides
  const std::vector<int64_t> x = {1, 2, 0, 0, 5, 3, 0, 4, 0, 0, 0, 7, 9, 3, 0, 0};
  std::vector<int64_t> xt(16);
  for (int h = 0; h < 4; h++) for (int w = 0; w < 4; w++) xt[w * 4 + h] = x[h * 4 + w];
  int64_t* dX = dev_copy(xt);
  int64_t* dK = dev_copy(std::vector<int64_t>{1, 1, 1, 1, 1, 0, 1, 0, 0});
  int64_t* dY = dev_copy(std::vector<int64_t>(16, -1));
  am_conv2d_desc d{1, 1, 4, 4, 1, 3, 3, 1, 1, 1, 1, 1, 1};
  const int64_t xs[4] = {16, 16, 1, 4};            // element (n, c, h, w) at h + 4*w
  amCheck(am_conv2d_forward_strided_i64(nullptr, &d, dX, xs, dK, nullptr, nullptr, 1, dY, nullptr, AM_ACT_NONE));
  EXPECT((host_copy(dY, 16) == std::vector<int64_t>{1, 8, 5, 0, 8, 11, 5, 4, 8, 17, 10, 11, 9, 12, 10, 7}));
  // backward with grad_output = ones given through a stride-0 broadcast view (all four strides 0)
  int64_t* dOne = dev_copy(std::vector<int64_t>{1});
  int64_t* dGi = dev_copy(std::vector<int64_t>(16, -1));
  int64_t* dGk = dev_copy(std::vector<int64_t>(9, -1));
  const int64_t zs[4] = {0, 0, 0, 0};
  amCheck(am_conv2d_backward_strided_i64(nullptr, &d, dX, xs, dK, nullptr, dOne, zs, dGi, nullptr, dGk, nullptr, nullptr, 1));
  // grad_kernel[kh][kw] = sum of the input window shifted by (kh-1, kw-1); total input sum = 34
  auto gk = host_copy(dGk, 9);
  EXPECT(gk[4] == 34);
  auto gi = host_copy(dGi, 16);
  EXPECT(gi[5] == 6);                               // interior pixel: all 6 non-zero taps of the kernel reach it
  // ---- single-process multi-GPU entries: two "ranks" on device 0
  int devs[2] = {0, 0};
  am_mg_ctx* ctx = nullptr;
  amCheck(am_mg_init(2, devs, &ctx));
  const int64_t M = 5, N = 4, K = 4;               // gemm.nim:420-451
  const std::vector<int64_t> A5 = {5, 6, 5, 8, 8, 2, 8, 8, 0, 5, 4, 0, 4, 0, 5, 6, 4, 5, 0, 3};
  const std::vector<int64_t> B4 = {5, 3, 6, 0, 5, 2, 3, 3, 8, 8, 2, 0, 7, 7, 0, 0};
  int64_t r0[2], rn[2];
  const int64_t* Al[2]; const int64_t* Bl[2]; int64_t* Cl[2];
  for (int g = 0; g < 2; g++) {
    amCheck(am_mg_rows(ctx, M, g, &r0[g], &rn[g]));
    Al[g] = dev_copy(std::vector<int64_t>(A5.begin() + r0[g] * K, A5.begin() + (r0[g] + rn[g]) * K));
    Bl[g] = dev_copy(B4);
    Cl[g] = dev_copy(std::vector<int64_t>(M * N, -1));
  }
  EXPECT(rn[0] + rn[1] == M);
  amCheck(am_mg_gemm_rowsharded_i64(ctx, M, N, K, 1, Al, K, Bl, N, Cl, N));
  amCheck(am_mg_synchronize(ctx));
  const std::vector<int64_t> want = {151, 123, 58, 18, 170, 148, 70, 6, 57, 42, 23, 15, 102, 94, 34, 0, 66, 43, 39, 15};
  EXPECT(host_copy(Cl[0], 20) == want);
  EXPECT(host_copy(Cl[1], 20) == want);
  amCheck(am_mg_destroy(ctx));
  for (void* p : {(void*)dA, (void*)dB, (void*)dC, (void*)dX, (void*)dK, (void*)dY, (void*)dOne, (void*)dGi, (void*)dGk,
                  (void*)Al[0], (void*)Al[1], (void*)Bl[0], (void*)Bl[1], (void*)Cl[0], (void*)Cl[1]}) cudaFree(p);
}

int main() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { std::printf("SKIP: no GPU\n"); return 77; }
  gemm_known_answers<float>();
  gemm_known_answers<double>();
  gemm_known_answers<int32_t>();
  gemm_known_answers<int64_t>();
  conv_known_answers<float>();
  conv_known_answers<double>();
  conv_known_answers<int32_t>();
  conv_known_answers<int64_t>();
  nn_known_answers<float>();
  nn_known_answers<double>();
  boundary_round2();
  am_shutdown();
  std::printf("OK %d checks\n", checks);
  return 0;
}
