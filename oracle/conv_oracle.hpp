// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of Arraymancer's im2col + GEMM convolution
// (/root/reference/src/arraymancer/nn_primitives/fallback/conv.nim):
//   conv.nim:18-54    im2col                     -> im2col()
//   conv.nim:56-79    col2im (scatter-add)       -> col2im()
//   conv.nim:81-106   im2colgemm_conv2d          -> conv2d_forward()
//   conv.nim:108-140  im2colgemm_conv2d_gradient -> conv2d_backward()
//   nnp_convolution.nim:91-94  grad_bias = grad_output.sum(3).sum(2).sum(0)
// The per-image GEMMs go through the laser restatement (laser_gemm.hpp) — in the
// reference floats would reach a third-party BLAS here; BASELINE.json
// defines the float oracle as the reference's own gemm_strided.
// `dilation` is not in the CPU reference (it only exists on the cuDNN boundary,
// nnp_conv2d_cudnn.nim:20-22); dilation == 1 reproduces conv.nim exactly.
#pragma once
#include <vector>
#include "laser_gemm.hpp"

namespace laser_oracle {

struct ConvDims {
  int64_t N, C, H, W;        // input [N,C,H,W]
  int64_t Cout, kH, kW;      // kernel [Cout,C,kH,kW]
  int64_t padH, padW, sH, sW, dH, dW;
  int64_t Ho() const { return (H + 2 * padH - (dH * (kH - 1) + 1)) / sH + 1; }
  int64_t Wo() const { return (W + 2 * padW - (dW * (kW - 1) + 1)) / sW + 1; }
  int64_t Kcol() const { return C * kH * kW; }
};

// conv.nim:18-54.  Row c of the column matrix <-> (ci, kh, kw); column h*Wo + w.
template <class T>
static void im2col(const T* idata, const ConvDims& d, T* odata) {
  const int64_t Ho = d.Ho(), Wo = d.Wo(), flat_col = Ho * Wo, flat = d.H * d.W;
  const int64_t channels_col = d.Kcol();
  for (int64_t c = 0; c < channels_col; c++) {
    const int64_t w_offset = (c % d.kW) * d.dW - d.padW;
    const int64_t h_offset = ((c / d.kW) % d.kH) * d.dH - d.padH;
    const int64_t c_offset = (c / d.kW) / d.kH;
    for (int64_t h = 0; h < Ho; h++) {
      const int64_t row = h_offset + h * d.sH;
      for (int64_t w = 0; w < Wo; w++) {
        const int64_t col = w_offset + w * d.sW;
        T v = T(0);
        if (row >= 0 && col >= 0 && row < d.H && col < d.W) v = idata[c_offset * flat + row * d.W + col];
        odata[c * flat_col + h * Wo + w] = v;
      }
    }
  }
}

// conv.nim:56-79.  Scatter-add, iteration order c -> h -> w; result zero-initialised.
template <class T>
static void col2im(const T* cols, const ConvDims& d, T* out /* [C,H,W] */) {
  const int64_t Ho = d.Ho(), Wo = d.Wo(), flat_col = Ho * Wo;
  const int64_t channels_col = d.Kcol();
  for (int64_t i = 0; i < d.C * d.H * d.W; i++) out[i] = T(0);
  for (int64_t c = 0; c < channels_col; c++) {
    const int64_t w_offset = (c % d.kW) * d.dW - d.padW;
    const int64_t h_offset = ((c / d.kW) % d.kH) * d.dH - d.padH;
    const int64_t c_offset = (c / d.kW) / d.kH;
    for (int64_t h = 0; h < Ho; h++) {
      const int64_t row = h_offset + h * d.sH;
      for (int64_t w = 0; w < Wo; w++) {
        const int64_t col = w_offset + w * d.sW;
        if (row < 0 || col < 0 || row >= d.H || col >= d.W) continue;
        T& dst = out[(c_offset * d.H + row) * d.W + col];
        dst = arith<T>::add(dst, cols[c * flat_col + h * Wo + w]);
      }
    }
  }
}

// conv.nim:81-106.  Serial over images; out[i] = kernel_col . im2col(in[i]); then += bias.
template <class T>
static void conv2d_forward(const T* input, const T* kernel, const T* bias /* nullable: rank-0 bias */,
                           T* output, const ConvDims& d, int variant, int threads) {
  const int64_t Ho = d.Ho(), Wo = d.Wo(), HW = Ho * Wo, Kc = d.Kcol();
  std::vector<T> input_col((size_t)(Kc * HW));
  for (int64_t i = 0; i < d.N; i++) {
    im2col(input + i * d.C * d.H * d.W, d, input_col.data());
    gemm_strided<T>(d.Cout, HW, Kc, T(1), kernel, Kc, 1, input_col.data(), HW, 1, T(0),
                    output + i * d.Cout * HW, HW, 1, variant, threads);
  }
  if (bias) {
    for (int64_t i = 0; i < d.N; i++)
      for (int64_t co = 0; co < d.Cout; co++) {
        T* o = output + (i * d.Cout + co) * HW;
        for (int64_t p = 0; p < HW; p++) o[p] = arith<T>::add(o[p], bias[co]);
      }
  }
}

// conv.nim:108-140 + nnp_convolution.nim:91-94.
template <class T>
static void conv2d_backward(const T* input, const T* kernel, const T* grad_output, T* grad_input,
                            T* grad_weight, T* grad_bias /* nullable */, const ConvDims& d, int variant,
                            int threads) {
  const int64_t Ho = d.Ho(), Wo = d.Wo(), HW = Ho * Wo, Kc = d.Kcol();
  if (grad_bias) {
    // sum(3) then sum(2) then sum(0): W first, then H, then the batch.
    std::vector<T> s3((size_t)(d.N * d.Cout * Ho)), s2((size_t)(d.N * d.Cout));
    for (int64_t r = 0; r < d.N * d.Cout * Ho; r++) {
      T acc = T(0);
      for (int64_t w = 0; w < Wo; w++) acc = arith<T>::add(acc, grad_output[r * Wo + w]);
      s3[(size_t)r] = acc;
    }
    for (int64_t r = 0; r < d.N * d.Cout; r++) {
      T acc = T(0);
      for (int64_t h = 0; h < Ho; h++) acc = arith<T>::add(acc, s3[(size_t)(r * Ho + h)]);
      s2[(size_t)r] = acc;
    }
    for (int64_t co = 0; co < d.Cout; co++) {
      T acc = T(0);
      for (int64_t n = 0; n < d.N; n++) acc = arith<T>::add(acc, s2[(size_t)(n * d.Cout + co)]);
      grad_bias[co] = acc;
    }
  }
  for (int64_t i = 0; i < d.Cout * Kc; i++) grad_weight[i] = T(0);
  std::vector<T> input_col((size_t)(Kc * HW)), gcol((size_t)(Kc * HW)), gw((size_t)(d.Cout * Kc));
  for (int64_t i = 0; i < d.N; i++) {
    const T* gout = grad_output + i * d.Cout * HW;              // [Cout, HW]
    // grad_input_col = kernel_col^T * grad_output_col   (A = transposed view: rs=1, cs=Kc)
    gemm_strided<T>(Kc, HW, d.Cout, T(1), kernel, 1, Kc, gout, HW, 1, T(0), gcol.data(), HW, 1, variant,
                    threads);
    im2col(input + i * d.C * d.H * d.W, d, input_col.data());
    col2im(gcol.data(), d, grad_input + i * d.C * d.H * d.W);
    // grad_weight += grad_output_col * input_col^T      (B = transposed view: rs=1, cs=HW)
    gemm_strided<T>(d.Cout, Kc, HW, T(1), gout, HW, 1, input_col.data(), 1, HW, T(0), gw.data(), Kc, 1,
                    variant, threads);
    for (int64_t j = 0; j < d.Cout * Kc; j++) grad_weight[j] = arith<T>::add(grad_weight[j], gw[(size_t)j]);
  }
}

}  // namespace laser_oracle
