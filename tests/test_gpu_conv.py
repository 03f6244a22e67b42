"""GPU parity tests of the fused implicit-GEMM conv2d forward / backward, through the C-ABI, against
the CPU oracle (restated im2col+GEMM, oracle/conv_oracle.hpp) and the reference's known answers."""
import numpy as np
import pytest

from tests.golden import known_answers as KA

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

NP = {"f32": np.float32, "f64": np.float64, "i32": np.int32, "i64": np.int64}
FWD_TOL = {"f32": 5e-6, "f64": 1e-13}      # same as the GEMM tolerances
BWD_TOL = {"f32": 1e-4, "f64": 1e-12}      # batch-order reduction differs from the serial reference loop


@pytest.fixture(scope="module")
def am():
    import arraymancer_b200 as am
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return am


def rel(g, r):
    g = np.asarray(g, np.float64); r = np.asarray(r, np.float64)
    return np.linalg.norm(g - r) / max(np.linalg.norm(r), 1e-300)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("dt", ["i32", "i64", "f32", "f64"])
def test_known_answers(am, dt):
    # tests/nn_primitives/test_nnp_convolution.nim:21-46 and :54-134 (cuDNN twin: test_nnp_convolution_cudnn.nim:19-39)
    for case in (KA.CONV_SIMPLE, KA.CONV_STRIDED):
        x = np.array(case["input"], NP[dt]); k = np.array(case["kernel"], NP[dt])
        b = np.array(case["bias"], NP[dt]).reshape(-1, 1, 1)
        want = np.array(case["target"], NP[dt])
        got = am.conv2d(dev(x), dev(k), dev(b), tuple(case["padding"]), tuple(case["stride"])).cpu().numpy()
        if dt.startswith("i"):
            assert np.array_equal(got, want), case["src"]
        else:
            assert np.abs(got - want).mean() <= 1e-7, case["src"]      # the reference's own criterion


def _numeric_grad(f, x, h=1e-5):
    g = np.zeros_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        old = x[i]
        x[i] = old + h; fp = f(x)
        x[i] = old - h; fm = f(x)
        x[i] = old
        g[i] = (fp - fm) / (2 * h)
    return g


def _mre(y, t):
    d = np.maximum(np.abs(y), np.abs(t))
    return np.where(d == 0, 0.0, np.abs(y - t) / np.where(d == 0, 1.0, d)).mean()


def test_gradcheck_like_reference(am, oracle):
    # test_nnp_convolution.nim:137-168: f32 analytic backward vs float64 numeric gradient, MRE < 1e-6
    c = KA.CONV_GRADCHECK
    rng = np.random.default_rng(0)
    x = rng.random(c["input_shape"], dtype=np.float32); k = rng.random(c["kernel_shape"], dtype=np.float32)
    b = rng.random(c["bias_shape"], dtype=np.float32)
    pad, st = tuple(c["padding"]), tuple(c["stride"])
    out = am.conv2d(dev(x), dev(k), dev(b), pad, st)
    gi, gw, gb = am.conv2d_backward(dev(x), dev(k), dev(b), pad, st, (1, 1), torch.ones_like(out))
    dx, dk, db = x.astype(np.float64), k.astype(np.float64), b.astype(np.float64)
    t_in = _numeric_grad(lambda v: oracle.conv2d(v, dk, db, pad, st).sum(), dx.copy())
    t_w = _numeric_grad(lambda v: oracle.conv2d(dx, v, db, pad, st).sum(), dk.copy())
    t_b = _numeric_grad(lambda v: oracle.conv2d(dx, dk, v, pad, st).sum(), db.copy())
    assert _mre(gb.cpu().numpy().astype(np.float64), t_b) < c["tol_mre"]
    assert _mre(gw.cpu().numpy().astype(np.float64), t_w) < c["tol_mre"]
    assert _mre(gi.cpu().numpy().astype(np.float64), t_in) < c["tol_mre"]


CASES = [  # input, kernel, padding, stride, dilation
    ((2, 3, 4, 5), (2, 3, 3, 3), (1, 1), (1, 1), (1, 1)),
    ((3, 1, 28, 28), (20, 1, 5, 5), (0, 0), (1, 1), (1, 1)),       # LeNet cv1
    ((2, 20, 12, 12), (50, 20, 5, 5), (0, 0), (1, 1), (1, 1)),     # LeNet cv2
    ((2, 4, 9, 8), (5, 4, 3, 2), (1, 2), (2, 1), (1, 1)),
    ((2, 3, 11, 10), (70, 3, 3, 3), (2, 2), (1, 1), (2, 2)),       # > 64 output channels, dilation
    ((1, 130, 6, 6), (6, 130, 3, 3), (1, 0), (2, 2), (2, 1)),      # > 128 input channels
    ((5, 2, 7, 7), (3, 2, 1, 1), (0, 0), (1, 1), (1, 1)),          # 1x1 kernel
    ((1, 1, 5, 5), (1, 1, 5, 5), (0, 0), (1, 1), (1, 1)),          # single output pixel
    ((3, 5, 40, 37), (9, 5, 3, 3), (1, 1), (1, 1), (1, 1)),        # several row bands / images per CTA
    ((2, 3, 9, 9), (4, 3, 3, 3), (3, 3), (1, 1), (1, 1)),          # padding larger than the kernel reach
    ((70, 2, 6, 6), (3, 2, 3, 3), (1, 1), (1, 1), (1, 1)),         # many small images per CTA
    ((5, 20, 12, 12), (50, 20, 5, 5), (0, 0), (1, 1), (1, 1)),     # cv2 with an odd batch: partial image group / slices
    ((3, 6, 8, 8), (64, 6, 3, 3), (1, 1), (1, 1), (1, 1)),         # 64 output channels: full tensor-core N, padded taps
    ((2, 40, 10, 10), (33, 40, 3, 3), (0, 0), (1, 1), (1, 1)),     # 360 GEMM rows: several 128-row chunks
    ((4, 8, 9, 9), (16, 8, 2, 2), (0, 1), (1, 2), (1, 1)),         # even kernel, mixed stride, one-sided reach
    # single input channel (conv_c1.cu fused kernels): 'same' 3x3, odd widths (scalar staging / stores), odd Cout,
    # padded 5x5, more images than resident CTAs would hold at once
    ((5, 1, 12, 12), (7, 1, 3, 3), (1, 1), (1, 1), (1, 1)),
    ((3, 1, 13, 11), (20, 1, 5, 5), (0, 0), (1, 1), (1, 1)),
    ((4, 1, 10, 9), (5, 1, 5, 5), (2, 2), (1, 1), (1, 1)),
    ((700, 1, 8, 8), (3, 1, 3, 3), (0, 0), (1, 1), (1, 1)),
    ((2, 1, 28, 28), (20, 1, 5, 5), (4, 4), (1, 1), (1, 1)),       # padding 4 = kW - 1: full correlation
]


@pytest.fixture(params=["auto", "tc", "direct", "gather"])
def conv_path(request, am):
    """Every kernel family over the whole matrix: AUTO (per-shape pick), the tcgen05 kernels wherever they fit (TC),
    the smem-staged direct SIMT kernels (DIRECT) and the generic gather kernels (GATHER, the fallback)."""
    from arraymancer_b200 import _capi
    _capi.set_conv_path({"auto": _capi.CONV_AUTO, "tc": _capi.CONV_TC, "direct": _capi.CONV_DIRECT,
                         "gather": _capi.CONV_GATHER}[request.param])
    yield request.param
    _capi.set_conv_path(_capi.CONV_AUTO)


@pytest.mark.parametrize("dt", ["f32", "f64", "i32", "i64"])
@pytest.mark.parametrize("ci", range(len(CASES)))
def test_forward_backward_vs_oracle(am, oracle, conv_path, dt, ci):
    xs, ks, pad, st, dil = CASES[ci]
    rng = np.random.default_rng(100 + ci)
    if dt.startswith("f"):
        x = rng.random(xs).astype(NP[dt]); k = (rng.random(ks) - 0.5).astype(NP[dt]); b = rng.random((ks[0], 1, 1)).astype(NP[dt])
    else:
        hi = 2**30 if dt == "i32" else 2**62          # big values: products wrap mod 2^n
        x = rng.integers(-hi, hi, xs).astype(NP[dt]); k = rng.integers(-hi, hi, ks).astype(NP[dt])
        b = rng.integers(-hi, hi, (ks[0], 1, 1)).astype(NP[dt])
    want = oracle.conv2d(x, k, b, pad, st, dil)
    X, K_, B_ = dev(x), dev(k), dev(b)
    got = am.conv2d(X, K_, B_, pad, st, dil).cpu().numpy()
    go = (rng.random(want.shape) * 2 - 1).astype(NP[dt]) if dt.startswith("f") else rng.integers(-9, 9, want.shape).astype(NP[dt])
    wgi, wgw, wgb = oracle.conv2d_backward(x, k, go, True, pad, st, dil)
    gi, gw, gb = am.conv2d_backward(X, K_, B_, pad, st, dil, dev(go))
    gi, gw, gb = gi.cpu().numpy(), gw.cpu().numpy(), gb.cpu().numpy()
    if dt.startswith("i"):
        assert np.array_equal(got, want) and np.array_equal(gi, wgi) and np.array_equal(gw, wgw) and np.array_equal(gb, wgb)
    else:
        assert rel(got, want) <= FWD_TOL[dt]
        assert rel(gi, wgi) <= BWD_TOL[dt] and rel(gw, wgw) <= BWD_TOL[dt] and rel(gb, wgb) <= BWD_TOL[dt]


def test_no_bias_and_partial_gradients(am, oracle):
    rng = np.random.default_rng(3)
    x = rng.random((2, 3, 6, 6)).astype(np.float32); k = rng.random((4, 3, 3, 3)).astype(np.float32)
    out = am.conv2d(dev(x), dev(k), None, (1, 1))                      # rank-0 bias in the reference
    assert rel(out.cpu().numpy(), oracle.conv2d(x, k, None, (1, 1))) <= 5e-6
    gi, gw, gb = am.conv2d_backward(dev(x), dev(k), None, (1, 1), (1, 1), (1, 1), torch.ones_like(out))
    assert gb is None
    wgi, wgw, _ = oracle.conv2d_backward(x, k, np.ones_like(out.cpu().numpy()), False, (1, 1))
    assert rel(gi.cpu().numpy(), wgi) <= 1e-4 and rel(gw.cpu().numpy(), wgw) <= 1e-4


def test_partial_gradients(am, oracle):
    """need_input_grad / need_kernel_grad = False -> NULL pointers on the C ABI, that gradient is skipped."""
    rng = np.random.default_rng(9)
    x = rng.random((3, 4, 10, 10)).astype(np.float32); k = (rng.random((6, 4, 5, 5)) - 0.5).astype(np.float32)
    b = rng.random((6, 1, 1)).astype(np.float32)
    go = (rng.random((3, 6, 6, 6)) - 0.5).astype(np.float32)
    wgi, wgw, wgb = oracle.conv2d_backward(x, k, go)
    gi, gw, gb = am.conv2d_backward(dev(x), dev(k), dev(b), (0, 0), (1, 1), (1, 1), dev(go), need_kernel_grad=False)
    assert gw is None and gb is None and rel(gi.cpu().numpy(), wgi) <= 1e-4
    gi, gw, gb = am.conv2d_backward(dev(x), dev(k), dev(b), (0, 0), (1, 1), (1, 1), dev(go), need_input_grad=False)
    assert gi is None and rel(gw.cpu().numpy(), wgw) <= 1e-4 and rel(gb.cpu().numpy(), wgb) <= 1e-4


def test_errors(am):
    x = torch.zeros((1, 3, 8, 8), device="cuda"); k = torch.zeros((2, 4, 3, 3), device="cuda")
    with pytest.raises(IndexError):
        am.conv2d(x, k, None)                                          # channel mismatch
    with pytest.raises(ValueError):
        am.conv2d(x[0], k, None)                                       # rank
    with pytest.raises(ValueError):
        am.conv2d(torch.zeros((1, 4, 2, 2), device="cuda"), k, None)   # kernel larger than the input


def test_c4_lenet_batch4096_properties(am, oracle):
    """BASELINE configs[3]: LeNet conv layers at batch 4096 (f32).  The oracle is serial over images
    (conv.nim:99), so it checks a slice of the batch; the rest is covered by size-independent properties:
    per-image independence, linearity in the input, and the adjoint identity <conv(x), g> = <x, dgrad(g)>."""
    g = torch.Generator(device="cuda"); g.manual_seed(2024)
    for xs, ks in [((4096, 1, 28, 28), (20, 1, 5, 5)), ((4096, 20, 12, 12), (50, 20, 5, 5))]:
        X = torch.rand(xs, device="cuda", generator=g)
        fan_in = ks[1] * ks[2] * ks[3]
        W = torch.randn(ks, device="cuda", generator=g) * (2.0 / fan_in) ** 0.5     # Kaiming (conv2D.nim:165-173)
        B = torch.rand((ks[0], 1, 1), device="cuda", generator=g)
        out = am.conv2d(X, W, B)
        # oracle on 16 images spread over the batch
        idx = torch.tensor([0, 1, 2, 3, 1000, 1001, 2047, 2048, 2049, 3000, 3500, 4000, 4092, 4093, 4094, 4095], device="cuda")
        want = oracle.conv2d(X[idx].cpu().numpy(), W.cpu().numpy(), B.cpu().numpy())
        assert rel(out[idx].cpu().numpy(), want) <= 5e-6
        # per-image independence: a sub-batch gives the identical bits
        assert torch.equal(am.conv2d(X[1024:1536].contiguous(), W, B), out[1024:1536])
        # backward: oracle on the same 16 images with grad_output = ones (test_nnp_convolution.nim:161) ...
        go = torch.ones_like(out)
        gi, gw, gb = am.conv2d_backward(X, W, B, (0, 0), (1, 1), (1, 1), go)
        wgi, wgw, wgb = oracle.conv2d_backward(X[idx].cpu().numpy(), W.cpu().numpy(), go[idx].cpu().numpy())
        assert rel(gi[idx].cpu().numpy(), wgi) <= 1e-4
        gi_s, gw_s, gb_s = am.conv2d_backward(X[idx].contiguous(), W, B, (0, 0), (1, 1), (1, 1), go[idx].contiguous())
        assert rel(gw_s.cpu().numpy(), wgw) <= 1e-4 and rel(gb_s.cpu().numpy(), wgb) <= 1e-4
        # ... and at full batch: grad_bias with ones = N*Ho*Wo exactly representable, adjoint identities in float64
        assert torch.allclose(gb.flatten(), torch.full((ks[0],), float(out.shape[0] * out.shape[2] * out.shape[3]), device="cuda"))
        G = torch.rand(out.shape, device="cuda", generator=g) * 2 - 1
        gi, gw, gb = am.conv2d_backward(X, W, B, (0, 0), (1, 1), (1, 1), G)
        conv_nb = (out - B.reshape(1, -1, 1, 1)).double()
        lhs = (conv_nb * G.double()).sum().item()                                   # <conv(x) - bias, g>
        scale = (conv_nb.abs() * G.double().abs()).sum().item()                      # sum of |terms|: f32 rounding bound
        assert abs(lhs - (X.double() * gi.double()).sum().item()) <= 2e-6 * scale    # = <x, dgrad(g)>
        assert abs(lhs - (W.double() * gw.double()).sum().item()) <= 2e-6 * scale    # = <w, wgrad(g)>
        ref_b = G.double().sum(dim=(0, 2, 3))
        bound = 2e-6 * G.double().abs().sum(dim=(0, 2, 3))
        assert bool(((gb.flatten().double() - ref_b).abs() <= bound).all())
