"""GPU parity tests added in round 2: staging modes of the SIMT GEMM (128-bit loads) over the whole layout matrix, the
forced tcgen05 path over strided / negative / zero-stride operands, K = 32768 against the oracle, batched GEMM,
strided conv boundary, host-buffer entries with padded C, tuning knobs."""
import ctypes

import numpy as np
import pytest

from tests.test_gpu_gemm import F32_TOL, F64_TOL, NP, dev, rand, rel_fro

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
TDT = {"f32": torch.float32, "f64": torch.float64, "i32": torch.int32, "i64": torch.int64}


@pytest.fixture(scope="module")
def am():
    import arraymancer_b200 as am
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return am


def _assert_same(dt, got, want, tag=""):
    if dt in ("i32", "i64"):
        assert np.array_equal(got, want), (tag, int((got != want).sum()))
    else:
        assert rel_fro(got, want) <= (F32_TOL if dt == "f32" else F64_TOL), (tag, rel_fro(got, want))


# ---------------------------------------------------------------- SIMT staging modes (scalar / 16-B cp.async / LDG.128)
@pytest.mark.parametrize("dt", ["i32", "i64", "f64", "f32"])
@pytest.mark.parametrize("vec", [1, 0])
def test_simt_layout_matrix_vector_and_scalar_staging(am, oracle, dt, vec):
    """All four (A, B) layout combinations x C row/column-major, at shapes with tails in every dimension (M, N, K not
    multiples of the vector width or the tile), with and without the vector staging modes: same bits / tolerance."""
    from arraymancer_b200 import _capi
    _capi.set_tuning("simt_vec_load", vec)
    am.set_f32_path(am.F32_SIMT); am.set_f64_path(am.F64_SIMT)
    try:
        for (M, N, K) in [(131, 77, 203), (128, 256, 64), (260, 132, 36), (5, 3, 2), (64, 64, 1000)]:
            a, b = rand((M, K), dt, 3), rand((K, N), dt, 4)
            want = oracle.matmul(a, b)
            A_r, B_r = dev(a), dev(b)
            A_c, B_c = dev(a.T.copy()).t(), dev(b.T.copy()).t()           # same logical matrices, column-major storage
            for A in (A_r, A_c):
                for B in (B_r, B_c):
                    for order in ("C", "F"):
                        C = torch.empty((M, N), dtype=TDT[dt], device="cuda")
                        if order == "F":
                            C = C.t().contiguous().t()
                        am.gemm_strided(1, A, B, 0, C)
                        _assert_same(dt, C.cpu().numpy(), want, (M, N, K, A.stride(), B.stride(), order))
        # aligned row pitch but views that start at an odd element: the vector modes must fall back to scalar staging
        a, b = rand((100, 104), dt, 5), rand((104, 96), dt, 6)
        A, B = dev(a)[:, 1:101], dev(b)[1:101, :]
        C = torch.empty((100, 96), dtype=TDT[dt], device="cuda")
        am.gemm_strided(1, A, B, 0, C)
        _assert_same(dt, C.cpu().numpy(), oracle.matmul(np.ascontiguousarray(a[:, 1:101]), np.ascontiguousarray(b[1:101])))
    finally:
        _capi.set_tuning("simt_vec_load", 1)
        am.set_f32_path(am.F32_AUTO); am.set_f64_path(am.F64_AUTO)


# ---------------------------------------------------------------- forced tcgen05 path x stride matrix (both split_pack kernels)
@pytest.mark.parametrize("scalar_pack", [0, 1])
def test_f32_tc_stride_variants(am, oracle, scalar_pack):
    """test_stride_variants at a shape the tensor-core path really takes, with F32_TC forced: step-2, transposed,
    negative (raw C ABI) and zero strides through split_pack_kernel (scalar) and split_pack_vec_kernel (vector),
    both k_fast branches."""
    from arraymancer_b200 import _capi
    am.set_f32_path(am.F32_TC)
    _capi.set_tuning("pack_scalar", scalar_pack)
    try:
        M, N, K = 300, 260, 520
        ap, bp = rand((2 * K, 2 * K), "f32", 11), rand((2 * K, 2 * K), "f32", 12)
        Ap, Bp = dev(ap), dev(bp)
        a_views = [lambda x: x[:M, :K], lambda x: x[:2 * M:2, ::2][:, :K], lambda x: x[:K, :M].T,
                   lambda x: (np.broadcast_to(x[0:1, :K], (M, K)) if isinstance(x, np.ndarray) else x[0:1, :K].expand(M, K))]
        b_views = [lambda x: x[:K, :N], lambda x: x[1::2, ::2][:K, :N], lambda x: x[:N, :K].T]
        for av in a_views:
            for bv in b_views:
                a_np, b_np = av(ap), bv(bp)
                for order in ("C", "F"):
                    c0 = rand((M, N), "f32", 13, False)
                    want = np.array(c0, order=order)
                    oracle.gemm_strided(2, a_np, b_np, 1, want)
                    C = dev(c0)
                    if order == "F":
                        C = C.t().contiguous().t()
                    am.gemm_strided(2, av(Ap), bv(Bp), 1, C)
                    assert rel_fro(C.cpu().numpy(), want) <= F32_TOL
        # negative strides through the raw C ABI: A bottom-up and right-to-left
        a, b = rand((M, K), "f32", 5), rand((K, N), "f32", 6)
        A, B = dev(a), dev(b)
        C = torch.zeros((M, N), device="cuda")
        lib = _capi.lib()
        _capi.check(lib.am_gemm_strided_f32(None, M, N, K, 1.0, A.data_ptr() + 4 * (M * K - 1), -K, -1, B.data_ptr(), N, 1,
                                            0.0, C.data_ptr(), N, 1))
        torch.cuda.synchronize()
        assert rel_fro(C.cpu().numpy(), oracle.matmul(np.ascontiguousarray(a[::-1, ::-1]), b)) <= F32_TOL
    finally:
        _capi.set_tuning("pack_scalar", 0)
        am.set_f32_path(am.F32_AUTO)


def test_f32_tc_k32768_vs_oracle(am, oracle):
    """The stated float32 tolerance is for K <= 32768: check the tensor-core path AT K = 32768 against the oracle
    (and the fp64 ground truth), on a shape small enough for the oracle: 512 x 512 x 32768."""
    M, N, K = 512, 512, 32768
    a, b = rand((M, K), "f32", 1234), rand((K, N), "f32", 1235)
    R = oracle.matmul(a, b)
    T = a.astype(np.float64) @ b.astype(np.float64)
    am.set_f32_path(am.F32_TC)
    try:
        for order in ("C", "F"):
            C = torch.empty((M, N), device="cuda")
            if order == "F":
                C = C.t().contiguous().t()
            am.gemm_strided(1, dev(a), dev(b), 0, C)
            G = C.cpu().numpy()
            assert rel_fro(G, R) <= F32_TOL, rel_fro(G, R)
            assert rel_fro(G, T) <= 3 * rel_fro(R, T) + 1e-7, (rel_fro(G, T), rel_fro(R, T))
    finally:
        am.set_f32_path(am.F32_AUTO)


def test_f64_32768_sampled_rows(am, oracle):
    """BASELINE configs[4], float64: DGEMM with K = 32768 (2048 x 2048 x 32768 keeps it to seconds), oracle on 16 rows."""
    M, N, K = 2048, 2048, 32768
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    A = torch.rand((M, K), device="cuda", dtype=torch.float64, generator=g) * 2 - 1
    B = torch.rand((K, N), device="cuda", dtype=torch.float64, generator=g) * 2 - 1
    C = torch.empty((M, N), device="cuda", dtype=torch.float64)
    am.gemm_strided(1, A, B, 0, C)
    rows = np.sort(np.random.default_rng(2).choice(M, 16, replace=False))
    rt = torch.from_numpy(rows).cuda()
    want = oracle.matmul(np.ascontiguousarray(A[rt].cpu().numpy()), B.cpu().numpy())
    assert rel_fro(C[rt].cpu().numpy(), want) <= F64_TOL


# ---------------------------------------------------------------- batched GEMM
@pytest.mark.parametrize("dt", ["f32", "f64", "i32", "i64"])
def test_gemm_strided_batched(am, oracle, dt):
    """am_gemm_strided_batched_* (cublas.nim:172-208): independent products in one launch, shared operand via batch
    stride 0, transposed views, alpha/beta, column-major C."""
    nb, M, N, K = 7, 45, 38, 61
    a, b = rand((nb, M, K), dt, 21), rand((nb, K, N), dt, 22)
    c0 = rand((nb, M, N), dt, 23, False)
    A, B = dev(a), dev(b)
    for alpha, beta in ((1, 0), (-2, 3)):
        C = dev(c0)
        am.gemm_strided_batched(alpha, A, B, beta, C)
        got = C.cpu().numpy()
        for i in range(nb):
            want = c0[i].copy()
            oracle.gemm_strided(alpha, a[i], b[i], beta, want)
            _assert_same(dt, got[i], want, ("batched", i, alpha, beta))
    # B shared by every product (batch stride 0), A transposed per batch, C column-major per batch
    at = dev(np.ascontiguousarray(a.transpose(0, 2, 1)))              # [nb, K, M]
    Cc = torch.empty((nb, N, M), dtype=TDT[dt], device="cuda").transpose(1, 2)
    am.gemm_strided_batched(1, at.transpose(1, 2), B[0:1].expand(nb, K, N), 0, Cc)
    got = Cc.cpu().numpy()
    for i in range(nb):
        _assert_same(dt, got[i], oracle.matmul(a[i], b[0]), ("batched shared B", i))
    # the conv-over-images case of fallback/conv.nim:99: one weight matrix times per-image column matrices
    w = rand((20, 25), dt, 31); cols = rand((64, 25, 576), dt, 32)
    O = torch.empty((64, 20, 576), dtype=TDT[dt], device="cuda")
    am.gemm_strided_batched(1, dev(w).unsqueeze(0).expand(64, 20, 25), dev(cols), 0, O)
    for i in (0, 17, 63):
        _assert_same(dt, O[i].cpu().numpy(), oracle.matmul(w, cols[i]), ("batched conv-like", i))


def test_gemm_strided_batched_large_products_take_tensor_cores(am, oracle):
    nb, n = 3, 1280
    a, b = rand((nb, n, n), "f32", 41), rand((nb, n, n), "f32", 42)
    C = torch.empty((nb, n, n), device="cuda")
    am.gemm_strided_batched(1, dev(a), dev(b), 0, C)
    for i in range(nb):
        assert rel_fro(C[i].cpu().numpy(), oracle.matmul(a[i], b[i])) <= F32_TOL


# ---------------------------------------------------------------- strided conv boundary
@pytest.mark.parametrize("dt", ["f32", "i64"])
def test_conv2d_strided_tensors(am, oracle, dt):
    """am_conv2d_*_strided_*: NHWC-stored input, Fortran-ordered grad_output, a sliced kernel — every tensor crosses
    the boundary with its own strides (cudnn.nim:59-75) and the results match the dense call / the oracle."""
    rng = np.random.default_rng(0)
    xs, ks = (3, 4, 9, 8), (6, 4, 3, 3)
    if dt == "f32":
        x = rng.random(xs).astype(np.float32); k = (rng.random(ks) - 0.5).astype(np.float32); b = rng.random((6, 1, 1)).astype(np.float32)
    else:
        x = rng.integers(-9, 9, xs).astype(np.int64); k = rng.integers(-9, 9, ks).astype(np.int64); b = rng.integers(-9, 9, (6, 1, 1)).astype(np.int64)
    pad, st, dil = (1, 1), (2, 1), (1, 1)
    want = oracle.conv2d(x, k, b, pad, st, dil)
    X_nhwc = dev(np.ascontiguousarray(x.transpose(0, 2, 3, 1))).permute(0, 3, 1, 2)        # NCHW view of NHWC storage
    Kbig = dev(np.concatenate([k, k], axis=1))[:, :4]                                        # channel slice of a wider kernel
    assert not X_nhwc.is_contiguous() and not Kbig.is_contiguous()
    out = am.conv2d(X_nhwc, Kbig, dev(b), pad, st, dil)
    _assert_same(dt, out.cpu().numpy(), want, "fwd strided")
    go = (rng.random(want.shape) * 2 - 1).astype(np.float32) if dt == "f32" else rng.integers(-5, 5, want.shape).astype(np.int64)
    wgi, wgw, wgb = oracle.conv2d_backward(x, k, go, True, pad, st, dil)
    GO_f = dev(np.ascontiguousarray(go.transpose(3, 2, 1, 0))).permute(3, 2, 1, 0)         # Fortran-ordered grad_output
    assert not GO_f.is_contiguous()
    gi, gw, gb = am.conv2d_backward(X_nhwc, Kbig, dev(b), pad, st, dil, GO_f)
    if dt == "f32":
        assert rel_fro(gi.cpu().numpy(), wgi) <= F32_TOL and rel_fro(gw.cpu().numpy(), wgw) <= 1e-4 and rel_fro(gb.cpu().numpy(), wgb) <= 1e-4
    else:
        assert np.array_equal(gi.cpu().numpy(), wgi) and np.array_equal(gw.cpu().numpy(), wgw) and np.array_equal(gb.cpu().numpy(), wgb)
    # raw entry: strided OUTPUTS too (grad_input written into an NHWC buffer, output into a padded buffer)
    from arraymancer_b200 import _capi
    lib = _capi.lib()
    suf = dt
    d = _capi.ConvDesc(3, 4, 9, 8, 6, 3, 3, 1, 1, 2, 1, 1, 1)
    N_, Co, Ho, Wo = want.shape
    obuf = torch.full((N_, Co, Ho, Wo + 3), -7, dtype=TDT[dt], device="cuda")
    oview = obuf[:, :, :, :Wo]
    s4 = lambda t: (ctypes.c_int64 * 4)(*t.stride())
    Xc, Kc, Bc = dev(x), dev(k), dev(b.reshape(-1))
    _capi.check(getattr(lib, f"am_conv2d_forward_strided_{suf}")(None, ctypes.byref(d), Xc.data_ptr(), None, Kc.data_ptr(), None,
                                                                  Bc.data_ptr(), 1, oview.data_ptr(), s4(oview), 0))
    torch.cuda.synchronize()
    _assert_same(dt, oview.cpu().numpy(), want, "fwd strided output")
    assert bool((obuf[:, :, :, Wo:] == -7).all())
    gibuf = torch.zeros((3, 9, 8, 4), dtype=TDT[dt], device="cuda")                          # NHWC storage for grad_input
    giv = gibuf.permute(0, 3, 1, 2)
    _capi.check(getattr(lib, f"am_conv2d_backward_strided_{suf}")(None, ctypes.byref(d), Xc.data_ptr(), None, Kc.data_ptr(), None,
                                                                   GO_f.data_ptr(), s4(GO_f), giv.data_ptr(), s4(giv), None, None, None, 1))
    torch.cuda.synchronize()
    _assert_same(dt, giv.cpu().numpy(), wgi, "grad_input strided output")


def test_conv2d_backward_null_input_is_rejected(am):
    from arraymancer_b200 import _capi
    lib = _capi.lib()
    d = _capi.ConvDesc(2, 3, 8, 8, 4, 3, 3, 0, 0, 1, 1, 1, 1)
    go = torch.ones((2, 4, 6, 6), device="cuda"); gb = torch.empty(4, device="cuda"); k = torch.ones((4, 3, 3, 3), device="cuda")
    rc = lib.am_conv2d_backward_f32(None, ctypes.byref(d), None, k.data_ptr(), go.data_ptr(), None, None, gb.data_ptr())
    assert rc == _capi.AM_ERR_INVALID and b"input" in lib.am_last_error()


# ---------------------------------------------------------------- host-buffer entries: padded C on the chunked path
@pytest.mark.parametrize("dt", ["f64", "i64", "f32"])
def test_host_entry_row_chunks_keep_gaps_of_padded_c(am, oracle, dt):
    """M >= 2048 takes the row-chunk schedule; with a padded host C (row pitch > N) the bytes
    between the rows must survive, and the rows must be right."""
    from arraymancer_b200 import _capi
    M, N, K = 2304, 200, 96
    a, b = rand((M, K), dt, 81), rand((K, N), dt, 82)
    sentinel = NP[dt](-5)
    c = np.full((M, N + 24), sentinel)
    cv = c[:, :N]
    ct = _capi.CTYPE[dt]
    it = c.itemsize
    _capi.check(getattr(_capi.lib(), f"am_host_gemm_strided_{dt}")(M, N, K, ct(1), a.ctypes.data, K, 1, b.ctypes.data, N, 1, ct(0),
                                                                    cv.ctypes.data, c.strides[0] // it, 1))
    assert np.all(c[:, N:] == sentinel), "gap columns of the padded host C were overwritten"
    _assert_same(dt, np.ascontiguousarray(cv), oracle.matmul(a, b), "host chunked")
    # column-strided C (csC = 2): takes the single-chunk path (C is uploaded first so the gaps survive)
    c2 = np.full((M, 2 * N), sentinel)
    cv2 = c2[:, ::2]
    _capi.check(getattr(_capi.lib(), f"am_host_gemm_strided_{dt}")(M, N, K, ct(1), a.ctypes.data, K, 1, b.ctypes.data, N, 1, ct(0),
                                                                    cv2.ctypes.data, c2.strides[0] // it, 2))
    assert np.all(c2[:, 1::2] == sentinel)
    _assert_same(dt, np.ascontiguousarray(cv2), oracle.matmul(a, b), "host csC=2")


def test_host_entry_f32_row_chunks_tensor_core_reuses_packed_b(am, oracle):
    """Row chunks of a float32 host product on the tcgen05 path: B is packed once into an explicit handle (chunk 0) and
    reused; the K-pipelined schedule is switched off by the tuning knob so this path is the one that runs."""
    from arraymancer_b200 import _capi
    M, N, K = 4608, 1024, 1024
    a, b = rand((M, K), "f32", 91), rand((K, N), "f32", 92)
    c = np.empty((M, N), np.float32)
    _capi.set_tuning("host_rowchunks", 1)
    try:
        l0 = _capi.kernel_launch_count()
        _capi.check(_capi.lib().am_host_gemm_strided_f32(M, N, K, 1.0, a.ctypes.data, K, 1, b.ctypes.data, N, 1, 0.0, c.ctypes.data, N, 1))
        launches = _capi.kernel_launch_count() - l0
    finally:
        _capi.set_tuning("host_rowchunks", 0)
    rows = [0, 1, 2303, 2304, M - 1]
    assert rel_fro(c[rows], oracle.matmul(np.ascontiguousarray(a[rows]), b)) <= F32_TOL
    nchunks = -(-M // 768)
    assert launches == 1 + 2 * nchunks, launches          # one pack of B + (pack A, GEMM) per chunk


def test_tuning_knobs(am):
    from arraymancer_b200 import _capi
    assert _capi.get_tuning("tc_flush_kb") == 2 and _capi.get_tuning("simt_vec_load") == 1
    _capi.set_tuning("tc_group", 4)
    assert _capi.get_tuning("tc_group") == 4
    _capi.set_tuning("tc_group", 8)
    with pytest.raises(_capi.AmError):
        _capi.set_tuning("no_such_knob", 1)


def test_matmul_k_zero_returns_zeros(am):
    a = am.cuda(np.zeros((4, 0), np.float32)); b = am.cuda(np.zeros((0, 3), np.float32))
    assert np.array_equal((a * b).cpu(), np.zeros((4, 3), np.float32))


# ---------------------------------------------------------------- single-process multi-GPU C entries (am_mg_*)
def _mg_devices():
    n = torch.cuda.device_count()
    return [[0], [0, 0], [0, 0, 0]] + ([list(range(n))] if n > 1 else [])


@pytest.mark.parametrize("dt", ["f32", "f64", "i64", "i32"])
def test_mg_gemm_rowsharded(am, oracle, dt):
    """am_mg_gemm_rowsharded_*: every GPU of the context ends with the whole C.  On a one-GPU box the context lists
    device 0 several times (the ranks then share the GPU): the row partition, the peer stores of the fused float32
    epilogue and the copy-engine pushes of the other dtypes run exactly as across GPUs."""
    from arraymancer_b200.multi_gpu import MgContext
    cur = torch.cuda.current_device()
    for devs in _mg_devices():
        ctx = MgContext(devs)
        for (M, N, K) in [(1100, 520, 300), (37, 29, 150)]:
            a, b = rand((M, K), dt, 61), rand((K, N), dt, 62)
            want = oracle.matmul(a, b)
            A_local, Bs, Cs = [], [], []
            for g, d in enumerate(devs):
                r0, n = ctx.rows(M, g)
                A_local.append(torch.from_numpy(np.ascontiguousarray(a[r0:r0 + n])).to(f"cuda:{d}"))
                Bs.append(torch.from_numpy(b).to(f"cuda:{d}"))
                Cs.append(torch.full((M, N), -7, dtype=TDT[dt], device=f"cuda:{d}"))
            assert sum(ctx.rows(M, g)[1] for g in range(len(devs))) == M
            ctx.gemm_rowsharded(1, A_local, Bs, Cs)
            ctx.synchronize()
            for g in range(len(devs)):
                _assert_same(dt, Cs[g].cpu().numpy(), want, ("mg", devs, g, M, N, K))
        ctx.close()
    assert torch.cuda.current_device() == cur          # the caller's device is restored


def test_mg_host_gemm_f32(am, oracle):
    from arraymancer_b200.multi_gpu import MgContext
    for devs in _mg_devices():
        ctx = MgContext(devs)
        M, N, K = 1300, 640, 520
        a, b = rand((M, K), "f32", 71), rand((K, N), "f32", 72)
        A, B = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
        C = torch.empty((M, N), dtype=torch.float32).pin_memory()
        ctx.host_gemm_f32(1.0, A, B, C)
        assert rel_fro(C.numpy(), oracle.matmul(a, b)) <= F32_TOL, devs
        ctx.close()


def test_f64_dmma_tma_kernel_matches_register_staged_kernel_and_is_deterministic(am):
    """The TMA-fed DMMA kernel (producer warp + 4-stage mbarrier ring) against the register-staged one on every
    operand layout, edge tiles and a K tail; repeated runs must be bit-identical (a stage overwritten early would show
    up as run-to-run differences at large K)."""
    from arraymancer_b200 import _capi
    am.set_f64_path(am.F64_DMMA)
    try:
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        for (M, N, K) in [(2048, 2304, 8192), (300, 260, 1001 * 2), (130, 4100, 70)]:
            Ar = torch.rand((M, K), device="cuda", dtype=torch.float64, generator=g) - 0.5
            Br = torch.rand((K, N), device="cuda", dtype=torch.float64, generator=g) - 0.5
            Ac = Ar.t().contiguous().t(); Bc = Br.t().contiguous().t()
            for A in (Ar, Ac):
                for B in (Br, Bc):
                    for order in ("C", "F"):
                        C0 = torch.empty((M, N), device="cuda", dtype=torch.float64)
                        C1 = torch.empty((M, N), device="cuda", dtype=torch.float64)
                        if order == "F":
                            C0 = C0.t().contiguous().t(); C1 = C1.t().contiguous().t()
                        _capi.set_tuning("dmma_tma", 0); am.gemm_strided(1, A, B, 0, C0)
                        _capi.set_tuning("dmma_tma", 1); am.gemm_strided(1, A, B, 0, C1)
                        assert rel_fro(C1.cpu().numpy(), C0.cpu().numpy()) <= 1e-14, (M, N, K, A.stride(), B.stride(), order)
        M, N, K = 2048, 2048, 16384
        A = torch.rand((M, K), device="cuda", dtype=torch.float64, generator=g) - 0.5
        B = torch.rand((K, N), device="cuda", dtype=torch.float64, generator=g) - 0.5
        C1 = torch.empty((M, N), device="cuda", dtype=torch.float64); C2 = torch.empty_like(C1)
        am.gemm_strided(1, A, B, 0, C1)
        for _ in range(6):
            am.gemm_strided(1, A, B, 0, C2)
            assert torch.equal(C1, C2)
    finally:
        _capi.set_tuning("dmma_tma", 1)
        am.set_f64_path(am.F64_AUTO)
