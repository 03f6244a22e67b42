"""Pins the CPU oracle (oracle/) against every known-answer vector the reference's own tests
hold for the dense-contraction path, then against independent numpy maths."""
import numpy as np
import pytest

from tests.golden import known_answers as KA
from tests.conftest import splitmix64

NP = {"f32": np.float32, "f64": np.float64, "i32": np.int32, "i64": np.int64}
VARIANTS = [0, 1, 2, 3]


@pytest.mark.parametrize("case", KA.GEMM, ids=[c["name"] for c in KA.GEMM])
@pytest.mark.parametrize("variant", VARIANTS)
def test_gemm_known_answers_all_dtypes(oracle, case, variant):
    # the reference runs the same literals as int, float32, float64 (test_operators_blas.nim:53,75,106)
    for dt in ("f32", "f64", "i32", "i64"):
        a = np.array(case["a"], dtype=NP[dt])
        b = np.array(case["b"], dtype=NP[dt])
        want = np.array(case["ab"], dtype=NP[dt])
        got = oracle.matmul(a, b, variant=variant)
        assert np.array_equal(got, want), (case["name"], dt, variant)


def test_gemm_transposed_views(oracle):
    t = KA.TRANSPOSE
    a, b = np.array(t["a"]), np.array(t["b"])
    at, bt = np.array(t["at"]), np.array(t["bt"])
    want = np.array(t["expected"])
    assert np.array_equal(oracle.matmul(at.T, b), want)
    assert np.array_equal(oracle.matmul(a, bt.T), want)
    assert np.array_equal(oracle.matmul(at.T, bt.T), want)


def test_gemm_colmajor_reversed_slice(oracle):
    t = KA.COLMAJOR_SLICE
    a = np.array(t["a"])
    eig = np.asfortranarray(np.array(t["eigvecs"]))
    view = eig[:, ::-1]                      # eigvecs[_, ^1..0|-1]: negative column stride
    assert view.strides[1] < 0
    got = oracle.matmul(a, view)
    assert np.abs(got - np.array(t["expected"])).mean() < t["tol_mae"]


@pytest.mark.parametrize("dt", ["i32", "i64", "f32", "f64"])
def test_conv_known_answers(oracle, dt):
    for case in (KA.CONV_SIMPLE, KA.CONV_STRIDED):
        x = np.array(case["input"], dtype=NP[dt])
        k = np.array(case["kernel"], dtype=NP[dt])
        b = np.array(case["bias"], dtype=NP[dt]).reshape(-1, 1, 1)
        want = np.array(case["target"], dtype=NP[dt])
        got = oracle.conv2d(x, k, b, tuple(case["padding"]), tuple(case["stride"]))
        assert got.shape == want.shape
        assert np.array_equal(got, want), case["src"]


def _numeric_grad(f, x, h=1e-5):
    # nn_primitives/nnp_numerical_gradient.nim:17-44 — central differences in float64
    g = np.zeros_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        old = x[i]
        x[i] = old + h
        fp = f(x)
        x[i] = old - h
        fm = f(x)
        x[i] = old
        g[i] = (fp - fm) / (2 * h)
    return g


def _mre(y, t):
    # ml/metrics/common_error_functions.nim mean_relative_error: |y-t| / max(|y|,|t|), 0 if both 0
    d = np.maximum(np.abs(y), np.abs(t))
    r = np.where(d == 0, 0.0, np.abs(y - t) / np.where(d == 0, 1.0, d))
    return r.mean()


def test_conv_gradcheck_like_reference(oracle):
    c = KA.CONV_GRADCHECK
    rng = np.random.default_rng(0)
    x = rng.random(c["input_shape"], dtype=np.float32)
    k = rng.random(c["kernel_shape"], dtype=np.float32)
    b = rng.random(c["bias_shape"], dtype=np.float32)
    pad, st = tuple(c["padding"]), tuple(c["stride"])
    out = oracle.conv2d(x, k, b, pad, st)
    gin, gw, gb = oracle.conv2d_backward(x, k, np.ones_like(out), True, pad, st)
    dx, dk, db = x.astype(np.float64), k.astype(np.float64), b.astype(np.float64)
    t_in = _numeric_grad(lambda v: oracle.conv2d(v, dk, db, pad, st).sum(), dx.copy())
    t_w = _numeric_grad(lambda v: oracle.conv2d(dx, v, db, pad, st).sum(), dk.copy())
    t_b = _numeric_grad(lambda v: oracle.conv2d(dx, dk, v, pad, st).sum(), db.copy())
    assert _mre(gb.astype(np.float64), t_b) < c["tol_mre"]
    assert _mre(gw.astype(np.float64), t_w) < c["tol_mre"]
    assert _mre(gin.astype(np.float64), t_in) < c["tol_mre"]


# ------------------------------------------------------------------ independent maths
def _rand_int(seed, shape, dt, full_range):
    raw = splitmix64(seed, int(np.prod(shape)))
    if full_range:
        return raw.view(np.int64).astype(NP[dt]).reshape(shape) if dt == "i64" else \
            (raw & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32).reshape(shape)
    return (raw % np.uint64(100)).astype(NP[dt]).reshape(shape)


def _wrap_matmul(a, b, dt):
    u = np.uint64 if dt == "i64" else np.uint32
    with np.errstate(over="ignore"):
        return (a.view(u) @ b.view(u)).view(NP[dt])


@pytest.mark.parametrize("dt", ["i32", "i64"])
@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 3, 5), (7, 13, 600), (97, 33, 129), (200, 17, 530), (5, 129, 2100)])
def test_int_gemm_wraps_mod_2n(oracle, dt, shape):
    M, N, K = shape
    a = _rand_int(7, (M, K), dt, True)
    b = _rand_int(8, (K, N), dt, True)
    want = _wrap_matmul(np.ascontiguousarray(a), np.ascontiguousarray(b), dt)
    for variant in VARIANTS:
        assert np.array_equal(oracle.matmul(a, b, variant=variant), want)


@pytest.mark.parametrize("dt", ["i64", "f64"])
def test_alpha_beta_and_strides(oracle, dt):
    rng = np.random.default_rng(3)
    M, N, K = 37, 29, 300   # K spans 2 kc-blocks for 8-byte types (kc = 256)
    def mk(shape):
        return rng.integers(-50, 50, size=shape).astype(NP[dt])
    a_parent, b_parent = mk((2 * M, 2 * K)), mk((K, N))
    views_a = [a_parent[::2, ::2], a_parent[:M, :K], np.asfortranarray(a_parent[:M, :K]),
               a_parent[::-2, :K][:M], np.broadcast_to(a_parent[0:1, :K], (M, K))]
    views_b = [b_parent, np.asfortranarray(b_parent), b_parent[::-1, ::-1]]
    for alpha, beta in [(1, 0), (-3, 0), (1, 1), (2, 2), (-3, 2)]:
        for va in views_a:
            for vb in views_b:
                c0 = mk((M, N))
                for order in ("C", "F"):
                    c = np.array(c0, order=order)
                    want = alpha * (va.astype(np.int64) @ vb.astype(np.int64)) + beta * c0.astype(np.int64)
                    oracle.gemm_strided(alpha, va, vb, beta, c)
                    assert np.array_equal(c.astype(np.int64), want)


def test_beta_zero_never_reads_c(oracle):
    a = np.ones((9, 300)); b = np.ones((300, 20))
    c = np.full((9, 20), np.nan)
    oracle.gemm_strided(1.0, a, b, 0.0, c)
    assert np.array_equal(c, np.full((9, 20), 300.0))


def test_k_zero_leaves_c_untouched(oracle):
    # gemm.nim:203 TODO: K == 0 never enters the pc loop, C untouched even with beta != 1
    a = np.zeros((4, 0)); b = np.zeros((0, 5)); c = np.full((4, 5), 7.0)
    oracle.gemm_strided(1.0, a, b, 3.0, c)
    assert np.array_equal(c, np.full((4, 5), 7.0))


@pytest.mark.parametrize("dt,tol", [("f32", 2e-6), ("f64", 1e-14)])
def test_float_gemm_close_to_high_precision(oracle, dt, tol):
    rng = np.random.default_rng(11)
    a = (rng.random((150, 1100)) * 2 - 1).astype(NP[dt])
    b = (rng.random((1100, 90)) * 2 - 1).astype(NP[dt])
    truth = a.astype(np.longdouble) @ b.astype(np.longdouble)
    got = oracle.matmul(a, b)
    rel = np.linalg.norm((got - truth).astype(np.float64)) / np.linalg.norm(truth.astype(np.float64))
    assert rel < tol
    # thread-count independence (Appendix A.6)
    assert np.array_equal(got, oracle.matmul(a, b, threads=1))


def test_float_kc_block_order_matches_emulation(oracle):
    # Appendix A.3/A.4: fma-sequential inside a kc block (512 for f32), blocks added in order.
    rng = np.random.default_rng(5)
    K = 1200
    a = (rng.random((3, K)) * 2 - 1).astype(np.float32)
    b = (rng.random((K, 2)) * 2 - 1).astype(np.float32)
    want = np.zeros((3, 2), np.float32)
    for i in range(3):
        for j in range(2):
            c = np.float32(0)
            for pc in range(0, K, 512):
                ab = np.float32(0)
                for k in range(pc, min(K, pc + 512)):
                    # exact fma: product of two f32 is exact in f64; one rounding to f32
                    ab = np.float32(np.float64(a[i, k]) * np.float64(b[k, j]) + np.float64(ab))
                c = np.float32(c + ab) if pc else ab
            want[i, j] = c
    got = oracle.matmul(a, b)
    assert np.array_equal(got, want)


def test_im2col_col2im_against_numpy(oracle):
    rng = np.random.default_rng(2)
    x = rng.integers(-9, 9, size=(3, 7, 6)).astype(np.int64)
    for (kh, kw), pad, st in [((3, 3), (1, 1), (1, 1)), ((2, 3), (0, 2), (2, 1)), ((5, 5), (2, 2), (3, 3))]:
        cols = oracle.im2col(x, (kh, kw), pad, st)
        xp = np.pad(x, ((0, 0), (pad[0],) * 2, (pad[1],) * 2))
        Ho = (7 + 2 * pad[0] - kh) // st[0] + 1
        Wo = (6 + 2 * pad[1] - kw) // st[1] + 1
        want = np.zeros((3 * kh * kw, Ho * Wo), np.int64)
        for c in range(3 * kh * kw):
            ci, r, s = c // (kh * kw), (c // kw) % kh, c % kw
            for h in range(Ho):
                for w in range(Wo):
                    want[c, h * Wo + w] = xp[ci, r + h * st[0], s + w * st[1]]
        assert np.array_equal(cols, want)
        # col2im is the adjoint of im2col: <im2col(x), y> == <x, col2im(y)>
        y = rng.integers(-9, 9, size=cols.shape).astype(np.int64)
        back = oracle.col2im(y, (3, 7, 6), (kh, kw), pad, st)
        assert (cols * y).sum() == (x * back).sum()


def test_conv_matches_torch_including_dilation(oracle):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(4)
    x = rng.random((3, 4, 9, 8)); k = rng.random((5, 4, 3, 2)); b = rng.random((5, 1, 1))
    for pad, st, dil in [((0, 0), (1, 1), (1, 1)), ((1, 2), (2, 1), (1, 1)), ((2, 2), (1, 1), (2, 2)), ((1, 0), (2, 2), (2, 1))]:
        want = torch.nn.functional.conv2d(torch.from_numpy(x), torch.from_numpy(k), torch.from_numpy(b.reshape(-1)),
                                          stride=st, padding=pad, dilation=dil).numpy()
        got = oracle.conv2d(x, k, b, pad, st, dil)
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
        xt = torch.from_numpy(x).requires_grad_(True); kt = torch.from_numpy(k).requires_grad_(True)
        bt = torch.from_numpy(b.reshape(-1)).requires_grad_(True)
        g = rng.random(want.shape)
        torch.nn.functional.conv2d(xt, kt, bt, stride=st, padding=pad, dilation=dil).backward(torch.from_numpy(g))
        gin, gw, gb = oracle.conv2d_backward(x, k, g, True, pad, st, dil)
        assert np.allclose(gin, xt.grad.numpy(), rtol=1e-11, atol=1e-11)
        assert np.allclose(gw, kt.grad.numpy(), rtol=1e-11, atol=1e-11)
        assert np.allclose(gb.reshape(-1), bt.grad.numpy(), rtol=1e-11, atol=1e-11)
