"""Known-answer vectors transcribed from the reference's own tests for the dense-contraction
path.  Each entry cites the reference file:line holding the literal values.
These pin the oracle (tests/test_oracle_golden.py) and, through the C-ABI, the CUDA path
(tests/test_gpu_golden.py).  `python tests/golden/known_answers.py` re-writes the JSON copy
(tests/golden/known_answers.json) that non-Python consumers (the C++ host test) read.
"""
import json
import os

# ---------------------------------------------------------------- GEMM: gemm_strided self-tests
# src/arraymancer/laser/primitives/matrix_multiplication/gemm.nim:317-567
GEMM = [
    dict(name="laser_f64_3x3_3x2_a", src="gemm.nim:317-341", dtype="f64",
         a=[[1.0, 2, 3], [1.0, 1, 1], [1.0, 1, 1]], b=[[1.0, 1], [1.0, 1], [1.0, 1]],
         ab=[[6.0, 6], [3.0, 3], [3.0, 3]]),
    dict(name="laser_f64_3x3_3x2_b", src="gemm.nim:343-367", dtype="f64",
         a=[[1.0, 2, 3], [4.0, 5, 6], [7.0, 8, 9]], b=[[1.0, 1], [1.0, 1], [1.0, 1]],
         ab=[[6.0, 6], [15.0, 15], [24.0, 24]]),
    dict(name="laser_f64_2x3_3x2", src="gemm.nim:369-392; tests/tensor/test_operators_blas.nim:24-39", dtype="f64",
         a=[[1.0, 2, 3], [4.0, 5, 6]], b=[[7.0, 8], [9.0, 10], [11.0, 12]], ab=[[58.0, 64], [139.0, 154]]),
    dict(name="int_MltN_negatives", src="gemm.nim:394-418; test_operators_blas.nim:41-53,386-396", dtype="i64",
         a=[[-2, -3, -1], [3, 0, 4]], b=[[1, 5, 2, -1], [-3, 0, 3, 4], [6, -2, 7, -4]],
         ab=[[1, -8, -20, -6], [27, 7, 34, -19]]),
    dict(name="int_5x4_4x4", src="gemm.nim:420-451; test_operators_blas.nim:57-75,400-417", dtype="i64",
         a=[[5, 6, 5, 8], [8, 2, 8, 8], [0, 5, 4, 0], [4, 0, 5, 6], [4, 5, 0, 3]],
         b=[[5, 3, 6, 0], [5, 2, 3, 3], [8, 8, 2, 0], [7, 7, 0, 0]],
         ab=[[151, 123, 58, 18], [170, 148, 70, 6], [57, 42, 23, 15], [102, 94, 34, 0], [66, 43, 39, 15]]),
    dict(name="int_2x8_8x2", src="gemm.nim:453-483", dtype="i64",
         a=[[2, 4, 3, 1, 3, 1, 3, 1], [4, 3, 2, 4, 1, 0, 0, 0]],
         b=[[2, 2], [2, 1], [0, 3], [0, 1], [0, 2], [4, 3], [3, 3], [2, 1]],
         ab=[[27, 37], [14, 23]]),
    dict(name="int_8x2_2x8", src="gemm.nim:485-520", dtype="i64",
         a=[[2, 1], [1, 3], [2, 1], [1, 0], [3, 4], [2, 4], [3, 1], [4, 0]],
         b=[[2, 2, 0, 4, 0, 0, 4, 2], [2, 1, 2, 1, 2, 4, 4, 1]],
         ab=[[6, 5, 2, 9, 2, 4, 12, 5], [8, 5, 6, 7, 6, 12, 16, 5], [6, 5, 2, 9, 2, 4, 12, 5],
             [2, 2, 0, 4, 0, 0, 4, 2], [14, 10, 8, 16, 8, 16, 28, 10], [12, 8, 8, 12, 8, 16, 24, 8],
             [8, 7, 2, 13, 2, 4, 16, 7], [8, 8, 0, 16, 0, 0, 16, 8]]),
    dict(name="int_8x8_8x8", src="gemm.nim:522-567; test_operators_blas.nim:80-106,421-446", dtype="i64",
         a=[[2, 4, 3, 1, 3, 1, 3, 1], [1, 2, 1, 1, 2, 0, 4, 3], [2, 0, 0, 3, 0, 4, 4, 1],
            [1, 1, 4, 0, 3, 1, 3, 0], [3, 4, 1, 1, 4, 2, 3, 4], [2, 4, 0, 2, 3, 3, 3, 4],
            [3, 0, 0, 3, 1, 4, 3, 1], [4, 3, 2, 4, 1, 0, 0, 0]],
         b=[[2, 2, 0, 4, 0, 0, 4, 2], [2, 0, 0, 1, 1, 1, 3, 1], [0, 2, 2, 0, 2, 2, 3, 3],
            [0, 0, 1, 0, 4, 2, 4, 1], [0, 0, 1, 3, 4, 2, 4, 2], [4, 3, 4, 1, 4, 4, 0, 3],
            [3, 3, 0, 2, 1, 2, 3, 3], [2, 1, 2, 1, 2, 4, 4, 1]],
         ab=[[27, 23, 16, 29, 35, 32, 58, 37], [24, 19, 11, 23, 26, 30, 49, 27],
             [34, 29, 21, 21, 34, 34, 36, 32], [17, 22, 15, 21, 28, 25, 40, 33],
             [39, 27, 23, 40, 45, 46, 72, 41], [41, 26, 25, 34, 47, 48, 65, 38],
             [33, 28, 22, 26, 37, 34, 41, 33], [14, 12, 9, 22, 27, 17, 51, 23]]),
    dict(name="int_2x3_3x2", src="tests/tensor/test_operators_blas.nim:370-381", dtype="i64",
         a=[[1, 2, 3], [4, 5, 6]], b=[[7, 8], [9, 10], [11, 12]], ab=[[58, 64], [139, 154]]),
]

# tests/tensor/test_operators_blas.nim:133-155 — all four transpose combinations, float64.
TRANSPOSE = dict(
    src="tests/tensor/test_operators_blas.nim:133-155",
    a=[[1.0, 2, 3], [4.0, 5, 6]], b=[[7.0, 8], [9.0, 10], [11.0, 12]],
    at=[[1.0, 4], [2.0, 5], [3.0, 6]], bt=[[7.0, 9, 11], [8.0, 10, 12]],
    expected=[[58.0, 64], [139.0, 154]])

# tests/tensor/test_operators_blas.nim:157-193 — A (row-major) times a column-reversed
# slice of a COLUMN-MAJOR 2x2 matrix (negative column stride), expected to 1e-9 MAE.
COLMAJOR_SLICE = dict(
    src="tests/tensor/test_operators_blas.nim:157-193",
    a=[[0.6899999999999999, 0.4900000000000002], [-1.31, -1.21], [0.3900000000000001, 0.9900000000000002],
       [0.08999999999999986, 0.2900000000000005], [1.29, 1.09], [0.4899999999999998, 0.7900000000000005],
       [0.1899999999999999, -0.3099999999999996], [-0.8100000000000001, -0.8099999999999996],
       [-0.3100000000000001, -0.3099999999999996], [-0.71, -1.01]],
    eigvecs=[[-0.735178655544408, 0.6778733985280118], [0.6778733985280118, 0.735178655544408]],
    expected=[[0.827970186, -0.175115307], [-1.77758033, 0.142857227], [0.992197494, 0.384374989],
              [0.274210416, 0.130417207], [1.67580142, -0.209498461], [0.912949103, 0.175282444],
              [-0.0991094375, -0.349824698], [-1.14457216, 0.0464172582], [-0.438046137, 0.0177646297],
              [-1.22382056, -0.162675287]],
    tol_mae=1e-9)

# ---------------------------------------------------------------- conv2d
# tests/nn_primitives/test_nnp_convolution.nim:21-46 (int exact; float32 MAE <= 1e-7);
# same case on the cuDNN boundary: tests/nn_primitives/test_nnp_convolution_cudnn.nim:19-39.
CONV_SIMPLE = dict(
    src="tests/nn_primitives/test_nnp_convolution.nim:21-46",
    input=[[[[1, 2, 0, 0], [5, 3, 0, 4], [0, 0, 0, 7], [9, 3, 0, 0]]]],
    kernel=[[[[1, 1, 1], [1, 1, 0], [1, 0, 0]]]],
    bias=[0], padding=[1, 1], stride=[1, 1],
    target=[[[[1, 8, 5, 0], [8, 11, 5, 4], [8, 17, 10, 11], [9, 12, 10, 7]]]])

# tests/nn_primitives/test_nnp_convolution.nim:54-134 (int exact and float32 exact).
CONV_STRIDED = dict(
    src="tests/nn_primitives/test_nnp_convolution.nim:54-134",
    input=[[[[2, 2, 0, 2, 1], [0, 1, 1, 0, 2], [1, 2, 1, 2, 1], [2, 2, 0, 0, 2], [2, 1, 1, 1, 2]],
            [[2, 0, 1, 1, 1], [2, 2, 0, 0, 2], [2, 2, 1, 0, 0], [1, 1, 2, 2, 0], [2, 1, 1, 1, 0]],
            [[0, 1, 2, 2, 0], [1, 1, 1, 1, 0], [2, 1, 2, 2, 0], [0, 2, 2, 2, 1], [0, 0, 2, 2, 1]]]],
    kernel=[[[[-1, -1, -1], [1, 0, 1], [0, -1, 0]],
             [[1, 0, -1], [1, -1, 1], [0, 1, 0]],
             [[0, 0, 1], [-1, -1, -1], [-1, 0, 0]]],
            [[[0, 1, 0], [1, -1, -1], [1, 1, -1]],
             [[-1, 0, 1], [-1, -1, 1], [1, 1, 0]],
             [[0, 1, 1], [-1, 1, -1], [-1, -1, 0]]]],
    bias=[1, 0], padding=[1, 1], stride=[2, 2],
    target=[[[[2, -2, 0], [-3, 2, -5], [-2, -1, 0]], [[-7, 1, 0], [3, -3, 2], [1, 3, -2]]]])

# tests/nn_primitives/test_nnp_convolution.nim:137-168 — fwd+bwd gradient check: analytic f32
# backward vs float64 central-difference numeric gradient, mean relative error < 1e-6;
# shapes input [2,3,4,5], kernel [2,3,3,3], bias [2,1,1], pad 1, stride 1, grad_output = ones.
CONV_GRADCHECK = dict(src="tests/nn_primitives/test_nnp_convolution.nim:137-168",
                      input_shape=[2, 3, 4, 5], kernel_shape=[2, 3, 3, 3], bias_shape=[2, 1, 1],
                      padding=[1, 1], stride=[1, 1], tol_mre=1e-6)

ALL = dict(GEMM=GEMM, TRANSPOSE=TRANSPOSE, COLMAJOR_SLICE=COLMAJOR_SLICE, CONV_SIMPLE=CONV_SIMPLE,
           CONV_STRIDED=CONV_STRIDED, CONV_GRADCHECK=CONV_GRADCHECK)

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "known_answers.json")
    with open(out, "w") as f:
        json.dump(ALL, f, indent=1)
    print("wrote", out)


# ---- the reference's own vectors for the operators around the contractions
# tests/nn_primitives/test_nnp_maxpool.nim:21-32  (kernel (2,2), padding (0,0), stride (2,2))
MAXPOOL = {
    "input": [[1, 1, 2, 4], [5, 6, 7, 8], [3, 2, 1, 0], [1, 2, 3, 4]],      # reshape(1,1,4,4)
    "kernel": (2, 2), "padding": (0, 0), "stride": (2, 2),
    "maxpooled": [6, 8, 3, 4],                                              # reshape(1,1,2,2)
    "max_indices": [5, 7, 8, 15],
}
# tests/nn_primitives/test_nnp_loss.nim:29-44  (`~=` is |a - b| <= 2e-5)
SOFTMAX_CE = {
    "predicted": [[-3.44, 1.16, -0.81, 3.91]],
    "sparse_truth": [3],
    "loss": 0.0709, "tol": 2e-5,
    "grad_mre_tol": 1e-6,           # analytic backward vs loss * numerical_gradient, :57-60
}
