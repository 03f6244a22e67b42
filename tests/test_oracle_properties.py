"""CPU: randomised property tests of the oracle (the checker must itself be right): the restated laser gemm_strided
against exact integer arithmetic for arbitrary strides / alpha / beta, and the restated im2col conv against a direct
quadruple-loop definition — shapes drawn like the reference's stability test (test_stability_openmp.nim:30-53)."""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

from oracle import laser_oracle as orc  # noqa: E402


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 40), st.integers(1, 40), st.integers(1, 70), st.integers(-3, 3), st.integers(-2, 2),
       st.sampled_from(["rr", "cr", "rc", "tt"]), st.integers(0, 2**31))
def test_gemm_int64_wraps_exactly(M, N, K, alpha, beta, layout, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(-2**62, 2**62, (M, K), dtype=np.int64)
    b = rng.integers(-2**62, 2**62, (K, N), dtype=np.int64)
    c = rng.integers(-2**62, 2**62, (M, N), dtype=np.int64)
    A = np.asfortranarray(a) if layout[0] in "ct" else a         # column-major storage = transposed-view strides
    B = np.asfortranarray(b) if layout[1] in "ct" else b
    want = (alpha * (a.astype(object) @ b.astype(object)) + beta * c.astype(object))
    want = np.vectorize(lambda v: ((int(v) + 2**63) % 2**64) - 2**63)(want).astype(np.int64)
    got = c.copy()
    orc.gemm_strided(alpha, A, B, beta, got)
    assert np.array_equal(got, want)


def _direct_conv(x, w, b, pad, stride):
    N, C, H, W = x.shape
    Co, _, kH, kW = w.shape
    Ho = (H + 2 * pad[0] - kH) // stride[0] + 1
    Wo = (W + 2 * pad[1] - kW) // stride[1] + 1
    xp = np.zeros((N, C, H + 2 * pad[0], W + 2 * pad[1]), dtype=x.dtype)
    xp[:, :, pad[0]:pad[0] + H, pad[1]:pad[1] + W] = x
    out = np.zeros((N, Co, Ho, Wo), dtype=x.dtype)
    for ho in range(Ho):
        for wo in range(Wo):
            patch = xp[:, :, ho * stride[0]:ho * stride[0] + kH, wo * stride[1]:wo * stride[1] + kW]
            out[:, :, ho, wo] = np.tensordot(patch, w, axes=([1, 2, 3], [1, 2, 3]))
    return out + b.reshape(1, Co, 1, 1)


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 3), st.integers(1, 4), st.integers(3, 9), st.integers(3, 9), st.integers(1, 5), st.integers(1, 3),
       st.integers(1, 3), st.integers(0, 2), st.integers(0, 2), st.integers(1, 2), st.integers(1, 2), st.integers(0, 2**31))
def test_conv_oracle_matches_direct_definition(N, C, H, W, Co, kH, kW, pH, pW, sH, sW, seed):
    if H + 2 * pH < kH or W + 2 * pW < kW:
        return
    rng = np.random.default_rng(seed)
    x = rng.integers(-9, 10, (N, C, H, W)).astype(np.int64)
    w = rng.integers(-9, 10, (Co, C, kH, kW)).astype(np.int64)
    b = rng.integers(-9, 10, (Co, 1, 1)).astype(np.int64)
    want = _direct_conv(x, w, b, (pH, pW), (sH, sW))
    got = orc.conv2d(x, w, b, (pH, pW), (sH, sW))
    assert np.array_equal(got, want)
    # gradients of sum(out * go): compare with the definition through linearity in x and w
    go = rng.integers(-3, 4, want.shape).astype(np.int64)
    gi, gw, gb = orc.conv2d_backward(x, w, go, True, (pH, pW), (sH, sW))
    assert np.array_equal(gb.reshape(-1), go.sum(axis=(0, 2, 3)))
    dx = rng.integers(-2, 3, x.shape).astype(np.int64)
    lhs = (_direct_conv(x + dx, w, b, (pH, pW), (sH, sW)) - want) * go          # linear in x: <go, conv(dx)> = <gi, dx>
    assert int(lhs.sum()) == int((gi * dx).sum())
    dw = rng.integers(-2, 3, w.shape).astype(np.int64)
    lhs = (_direct_conv(x, w + dw, b, (pH, pW), (sH, sW)) - want) * go
    assert int(lhs.sum()) == int((gw * dw).sum())
