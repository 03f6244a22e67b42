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
  am_shutdown();
  std::printf("OK %d checks\n", checks);
  return 0;
}
