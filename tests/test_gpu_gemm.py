"""GPU parity tests of the GEMM path, through the C-ABI, against the CPU oracle (oracle/) and the
reference's known-answer vectors.  Integer results must be bit-exact; float tolerances are the ones
SURVEY.md §8(d) states (written next to each check)."""
import numpy as np
import pytest

from tests.conftest import splitmix64
from tests.golden import known_answers as KA

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

NP = {"f32": np.float32, "f64": np.float64, "i32": np.int32, "i64": np.int64}
F32_TOL = 5e-6      # ||G-R||_F/||R||_F for float32, K <= 32768
F64_TOL = 1e-13     # same for float64


@pytest.fixture(scope="module")
def am():
    import arraymancer_b200 as am
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return am


def rel_fro(g, r):
    g = np.asarray(g, dtype=np.float64); r = np.asarray(r, dtype=np.float64)
    return np.linalg.norm(g - r) / max(np.linalg.norm(r), 1e-300)


def rand(shape, dt, seed, full_range=True):
    raw = splitmix64(seed, int(np.prod(shape)))
    if dt == "i64":
        return (raw.view(np.int64) if full_range else (raw % np.uint64(100)).astype(np.int64)).reshape(shape)
    if dt == "i32":
        v = (raw & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32) if full_range else (raw % np.uint64(100)).astype(np.int32)
        return v.reshape(shape)
    u = (raw >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))      # U[0,1)
    return (u * 2 - 1).astype(NP[dt]).reshape(shape)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def check(am, oracle, dt, a, b, alpha=1, beta=0, c0=None, a_view=None, b_view=None, c_order="C"):
    """Run C <- alpha*A*B + beta*C on views built by the *_view callables (applied to both the numpy
    array and the device tensor so strides match) and compare with the oracle on the same views."""
    M, N = a.shape[0], b.shape[1]
    if c0 is None:
        c0 = rand((M, N), dt, 99, full_range=False)
    want = np.array(c0, order=c_order)
    oracle.gemm_strided(alpha, a, b, beta, want)
    A, B = dev(np.array(a)), dev(np.array(b))
    # rebuild identical strides on the device: copy the parent buffer, then as_strided
    C = dev(c0)
    if c_order == "F":
        C = C.t().contiguous().t()
    am.gemm_strided(alpha, A, B, beta, C)
    got = C.cpu().numpy()
    if dt in ("i32", "i64"):
        assert np.array_equal(got, want), f"{dt} not bit-exact: {(got != want).sum()} mismatches"
    else:
        tol = F32_TOL if dt == "f32" else F64_TOL
        assert rel_fro(got, want) <= tol, (dt, rel_fro(got, want))


# ------------------------------------------------------------------ the reference's known answers
@pytest.mark.parametrize("case", KA.GEMM, ids=[c["name"] for c in KA.GEMM])
def test_known_answers_cudatensor_star(am, case):
    # mirrors tests/tensor/test_operators_blas_cuda.nim:20-98: `a.cuda * b.cuda` then `.cpu`
    for dt in ("f32", "f64", "i32", "i64"):
        a = np.array(case["a"], dtype=NP[dt]); b = np.array(case["b"], dtype=NP[dt])
        got = (am.cuda(a) * am.cuda(b)).cpu()
        assert np.array_equal(got, np.array(case["ab"], dtype=NP[dt])), (case["name"], dt)


def test_transposed_views(am):
    t = KA.TRANSPOSE
    at, bt = am.cuda(np.array(t["at"])), am.cuda(np.array(t["bt"]))
    a, b = am.cuda(np.array(t["a"])), am.cuda(np.array(t["b"]))
    want = np.array(t["expected"])
    assert np.array_equal((at.transpose() * b).cpu(), want)
    assert np.array_equal((a * bt.transpose()).cpu(), want)
    assert np.array_equal((at.transpose() * bt.transpose()).cpu(), want)


def test_colmajor_reversed_slice(am):
    t = KA.COLMAJOR_SLICE
    a = am.cuda(np.array(t["a"]))
    eig = am.cuda(np.array(t["eigvecs"]))           # column-major on the device, like newMatrixUninitColMajor
    val = a * eig[:, ::-1]                          # eigvecs[_, ^1..0|-1]: negative column stride
    assert np.abs(val.cpu() - np.array(t["expected"])).mean() < t["tol_mae"]


def test_matrix_vector(am):
    # test_operators_blas.nim:118-131 / :320-331
    d = am.cuda(np.array([[1.0, -1, 2], [0.0, -3, 1]])); e = am.cuda(np.array([2.0, 1, 0]))
    assert np.array_equal((d * e).cpu(), np.array([1.0, -3.0]))
    bi = am.cuda(np.array([[-87, 44, 13, 1], [8, -16, 8, 91], [6, -2, 56, -56], [82, 70, 34, 23], [52, -70, 0, 53],
                           [35, 94, 39, 36]], dtype=np.int64))
    u = am.cuda(np.array([-91, 81, 69, -75], dtype=np.int64))
    assert np.array_equal((bi * u).cpu(), np.array([12303, -8297, 7356, -1171, -14377, 4420]))


def test_error_conventions(am):
    a = am.cuda(np.zeros((2, 3), np.float32))
    with pytest.raises(IndexError):
        a * a                                        # check_matmat -> IndexDefect
    with pytest.raises(ValueError):
        am.cuda(np.zeros(3, np.float32)) * a        # vector * matrix is not defined (operators_blas_l2l3_cuda.nim:87)


# ------------------------------------------------------------------ randomized parity
SHAPES = [(1, 1, 1), (2, 3, 5), (13, 8, 63), (64, 65, 127), (129, 191, 193), (255, 257, 64), (511, 33, 513), (100, 300, 17)]


@pytest.mark.parametrize("dt", ["i64", "i32", "f64", "f32"])
@pytest.mark.parametrize("shape", SHAPES)
def test_random_shapes(am, oracle, dt, shape):
    M, N, K = shape
    check(am, oracle, dt, rand((M, K), dt, 7), rand((K, N), dt, 8))
    check(am, oracle, dt, rand((M, K), dt, 7), rand((K, N), dt, 8), c_order="F")


@pytest.mark.parametrize("dt", ["i64", "i32", "f64", "f32"])
@pytest.mark.parametrize("alpha,beta", [(1, 1), (-3, 0), (-3, 2), (1, 2)])
def test_alpha_beta(am, oracle, dt, alpha, beta):
    full = dt in ("i32", "i64")
    check(am, oracle, dt, rand((70, 300), dt, 1, full), rand((300, 90), dt, 2, full), alpha, beta,
          c0=rand((70, 90), dt, 3, full))


@pytest.mark.parametrize("dt", ["i64", "f64", "i32", "f32"])
def test_stride_variants(am, oracle, dt):
    """row-major, col-major, transposed, step-2 slices, negative steps, stride-0 broadcast — every
    (rowStride, colStride) pattern the reference's callers produce."""
    M, N, K = 37, 29, 150
    ap, bp = rand((2 * K, 2 * K), dt, 11), rand((2 * K, 2 * K), dt, 12)      # parents big enough for every view
    Ap, Bp = dev(ap), dev(bp)
    a_views = [lambda x: x[:M, :K], lambda x: x[:2 * M:2, ::2], lambda x: x[:K, :M].T, lambda x: x[M - 1::-1, :K] if isinstance(x, np.ndarray) else x[:M, :K].flip(0),
               lambda x: (np.broadcast_to(x[0:1, :K], (M, K)) if isinstance(x, np.ndarray) else x[0:1, :K].expand(M, K))]
    b_views = [lambda x: x[:K, :N], lambda x: x[1::2, ::2][:K, :N], lambda x: x[:N, :K].T]
    for ai, av in enumerate(a_views):
        for bv in b_views:
            a_np, b_np = av(ap), bv(bp)
            A_t, B_t = av(Ap), bv(Bp)
            if ai == 3:      # torch has no negative strides: build the same logical matrix by flipping (copy)
                a_np = np.ascontiguousarray(a_np)
            for order in ("C", "F"):
                c0 = rand((M, N), dt, 13, False)
                want = np.array(c0, order=order)
                oracle.gemm_strided(2, a_np, b_np, 1, want)
                C = dev(c0)
                if order == "F":
                    C = C.t().contiguous().t()
                am.gemm_strided(2, A_t, B_t, 1, C)
                got = C.cpu().numpy()
                if dt in ("i32", "i64"):
                    assert np.array_equal(got, want)
                else:
                    assert rel_fro(got, want) <= (F32_TOL if dt == "f32" else F64_TOL)


@pytest.mark.parametrize("path", ["dmma", "simt"])
def test_f64_both_kernels(am, oracle, path):
    """float64 has two kernels (DMMA tensor pipe / DFMA SIMT); force each over odd shapes, layouts, alpha/beta."""
    am.set_f64_path(am.F64_DMMA if path == "dmma" else am.F64_SIMT)
    try:
        for (M, N, K) in [(1, 1, 1), (7, 9, 5), (129, 130, 67), (300, 257, 1001), (1024, 1024, 512)]:
            a, b = rand((M, K), "f64", 51), rand((K, N), "f64", 52)
            check(am, oracle, "f64", a, b)
            check(am, oracle, "f64", a, b, alpha=-3, beta=2, c0=rand((M, N), "f64", 53), c_order="F")
        # transposed operand views
        a, b = rand((200, 300), "f64", 54), rand((150, 300), "f64", 55)
        C = torch.empty((200, 150), dtype=torch.float64, device="cuda")
        am.gemm_strided(1, dev(a), dev(b).t(), 0, C)
        assert rel_fro(C.cpu().numpy(), oracle.matmul(a, b.T)) <= F64_TOL
    finally:
        am.set_f64_path(am.F64_AUTO)


def test_negative_strides_through_raw_capi(am, oracle):
    """torch cannot express negative strides, the C ABI can: call it with raw pointers."""
    import ctypes
    from arraymancer_b200 import _capi
    M, N, K = 33, 21, 50
    a, b = rand((M, K), "i64", 5), rand((K, N), "i64", 6)
    A, B = dev(a), dev(b)
    C = torch.zeros((M, N), dtype=torch.int64, device="cuda")
    lib = _capi.lib()
    # A viewed bottom-up and right-to-left: pointer to the last element, strides (-K, -1)
    pa = A.data_ptr() + 8 * (M * K - 1)
    _capi.check(lib.am_gemm_strided_i64(None, M, N, K, 1, pa, -K, -1, B.data_ptr(), N, 1, 0, C.data_ptr(), N, 1))
    torch.cuda.synchronize()
    assert np.array_equal(C.cpu().numpy(), oracle.matmul(a[::-1, ::-1], b))
    # stride-0 broadcast row (p_shapeshifting.nim:76-84)
    _capi.check(lib.am_gemm_strided_i64(None, M, N, K, 1, A.data_ptr(), 0, 1, B.data_ptr(), N, 1, 0, C.data_ptr(), N, 1))
    torch.cuda.synchronize()
    assert np.array_equal(C.cpu().numpy(), oracle.matmul(np.broadcast_to(a[0:1], (M, K)), b))


def test_beta_zero_never_reads_c_and_k_zero_leaves_c(am):
    for dt in (torch.float32, torch.float64):
        A = torch.ones((40, 70), dtype=dt, device="cuda"); B = torch.ones((70, 50), dtype=dt, device="cuda")
        C = torch.full((40, 50), float("nan"), dtype=dt, device="cuda")
        am.gemm_strided(1, A, B, 0, C)
        assert torch.equal(C, torch.full_like(C, 70.0))
        C.fill_(7.0)
        am.gemm_strided(1, A[:, :0], B[:0, :], 3, C)      # K == 0: untouched (gemm.nim:203)
        assert torch.equal(C, torch.full_like(C, 7.0))


def test_stability_like_reference_openmp_stress(am, oracle):
    # tests/stability_tests/test_stability_openmp.nim:30-53: 100 random-shape (2..100) integer matmuls
    rng = np.random.default_rng(1337)
    for _ in range(100):
        M, K, N = (int(v) for v in rng.integers(2, 101, size=3))
        a = rng.integers(-100, 100, size=(M, K)).astype(np.int64); b = rng.integers(-100, 100, size=(K, N)).astype(np.int64)
        assert np.array_equal((am.cuda(a) * am.cuda(b)).cpu(), a @ b)


# ------------------------------------------------------------------ BASELINE configs
def test_c1_int64_1500_bit_exact(am, oracle):
    # benchmarks/integer_matmul.nim:11-12: int64 1500x1500, values U{0..99}; splitmix64 seeds 42 / 43
    a, b = rand((1500, 1500), "i64", 42, False), rand((1500, 1500), "i64", 43, False)
    got = (am.cuda(a) * am.cuda(b)).cpu()
    assert np.array_equal(got, oracle.matmul(a, b))
    # full-range operands exercise the wrap mod 2^64
    a, b = rand((1500, 1500), "i64", 7), rand((1500, 1500), "i64", 8)
    assert np.array_equal((am.cuda(a) * am.cuda(b)).cpu(), oracle.matmul(a, b))


def kostya(n):
    # benchmarks/kostya_matmul.nim:6-11
    i = np.arange(n, dtype=np.float64)[:, None]; j = np.arange(n, dtype=np.float64)[None, :]
    return (i - j) * (i + j) / n / n


def test_c2_float64_kostya(am, oracle):
    n = 1500
    a = kostya(n)
    got = (am.cuda(a) * am.cuda(a)).cpu()
    assert rel_fro(got, oracle.matmul(a, a)) <= F64_TOL
    # full size 8192: the oracle on 64 sampled rows (full K), stated in SURVEY §8d
    n = 8192
    a = kostya(n)
    A = dev(a); C = torch.empty((n, n), dtype=torch.float64, device="cuda")
    am.gemm_strided(1, A, A, 0, C)
    rows = np.random.default_rng(0).choice(n, 64, replace=False)
    want = oracle.matmul(np.ascontiguousarray(a[rows]), a)
    assert rel_fro(C[torch.from_numpy(rows).cuda()].cpu().numpy(), want) <= F64_TOL
    # size-independent property: symmetry of A*A^T at full size
    am.gemm_strided(1, A, A.t(), 0, C)
    assert torch.allclose(C, C.t(), rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("path", ["tc", "tc_1cta", "simt"])
def test_c3_float32_3xtf32_accuracy(am, oracle, path):
    am.set_f32_path({"tc": am.F32_TC, "tc_1cta": am.F32_TC_1CTA, "simt": am.F32_SIMT}[path])
    try:
        for (M, N, K) in [(512, 768, 1024), (300, 700, 100), (1024, 512, 16384)]:
            a, b = rand((M, K), "f32", 1234), rand((K, N), "f32", 1235)
            R = oracle.matmul(a, b)
            T = a.astype(np.float64) @ b.astype(np.float64)
            for c_order in ("C", "F"):
                C = torch.empty((M, N), dtype=torch.float32, device="cuda")
                if c_order == "F":
                    C = C.t().contiguous().t()
                am.gemm_strided(1, dev(a), dev(b), 0, C)
                G = C.cpu().numpy()
                assert rel_fro(G, R) <= F32_TOL, (path, M, N, K, rel_fro(G, R))
                # err(G,T) <= 3*err(R,T) + 1e-7  (a plain 1xTF32 kernel sits near 1e-3 and fails this)
                assert rel_fro(G, T) <= 3 * rel_fro(R, T) + 1e-7, (path, rel_fro(G, T), rel_fro(R, T))
    finally:
        am.set_f32_path(am.F32_AUTO)


def test_c3_float32_16384_views_sampled(am, oracle):
    """16384^2 with transposed / stepped views; oracle on 32 sampled rows (full K) — SURVEY §8d."""
    n = 16384
    g = torch.Generator(device="cuda"); g.manual_seed(1234)
    P = torch.rand((n, n), device="cuda", generator=g) * 2 - 1
    Q = torch.rand((n, n), device="cuda", generator=g) * 2 - 1
    C = torch.empty((n, n), device="cuda")
    rows = np.sort(np.random.default_rng(1).choice(n, 32, replace=False))
    rt = torch.from_numpy(rows).cuda()
    for name, A, B in [("rowmajor", P, Q), ("A^T view", P.t(), Q), ("B^T view", P, Q.t())]:
        am.gemm_strided(1, A, B, 0, C)
        want = oracle.matmul(np.ascontiguousarray(A[rt].cpu().numpy()), B.cpu().numpy())
        assert rel_fro(C[rt].cpu().numpy(), want) <= F32_TOL, name
    # column-major C (the CudaTensor default)
    Cf = torch.empty((n, n), device="cuda").t()
    am.gemm_strided(1, P, Q, 0, Cf)
    am.gemm_strided(1, P, Q, 0, C)
    # the operand roles (TMEM lanes vs columns) swap with C's layout, so the roundings differ: not bit-equal
    assert rel_fro(Cf[rt].cpu().numpy(), C[rt].cpu().numpy()) <= 2e-6


def test_integer_linearity_at_full_size(am):
    """size-independent property at 4096^2: (A1+A2)*B == A1*B + A2*B mod 2^64."""
    n = 4096
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    mk = lambda: torch.randint(-2**62, 2**62, (n, n), device="cuda", dtype=torch.int64, generator=g)
    A1, A2, B = mk(), mk(), mk()
    C1 = torch.empty_like(A1); C2 = torch.empty_like(A1); C3 = torch.empty_like(A1)
    am.gemm_strided(1, A1, B, 0, C1); am.gemm_strided(1, A2, B, 0, C2); am.gemm_strided(1, A1 + A2, B, 0, C3)
    assert torch.equal(C1 + C2, C3)
    am.gemm_strided(1, A2, B, 1, C1)               # beta = 1 accumulates in place
    assert torch.equal(C1, C3)


# ------------------------------------------------------------------ other boundary entries
def test_cublas_shaped_adapter(am, oracle):
    # cublas_gemm (cublas.nim:142-170): column-major buffers, op N / op T
    m, n, k = 70, 50, 90
    for dt, tdt in (("f32", torch.float32), ("f64", torch.float64)):
        a, b = rand((m, k), dt, 21), rand((k, n), dt, 22)
        want = oracle.matmul(a, b)
        for ta in (0, 1):
            for tb in (0, 1):
                Abuf = dev(a.T if ta == 0 else a).reshape(-1)      # op N: col-major A (= a^T rows); op T: row-major a
                Bbuf = dev(b.T if tb == 0 else b).reshape(-1)
                Cbuf = torch.empty(m * n, dtype=tdt, device="cuda")
                am.cublas_gemm(ta, tb, m, n, k, 1, Abuf, m if ta == 0 else k, Bbuf, k if tb == 0 else n, 0, Cbuf, m)
                got = Cbuf.reshape(n, m).t().cpu().numpy()
                assert rel_fro(got, want) <= (F32_TOL if dt == "f32" else F64_TOL)


def test_host_buffer_entry(am, oracle):
    import ctypes
    from arraymancer_b200 import _capi
    lib = _capi.lib()
    for dt in ("i64", "f64", "f32", "i32"):
        a, b = rand((90, 130), dt, 31), rand((130, 70), dt, 32)
        c = np.zeros((90, 70), NP[dt])
        ct = _capi.CTYPE[dt]
        fn = getattr(lib, f"am_host_gemm_strided_{dt}")
        _capi.check(fn(90, 70, 130, ct(1), a.ctypes.data, 130, 1, b.ctypes.data, 70, 1, ct(0), c.ctypes.data, 70, 1))
        want = oracle.matmul(a, b)
        if dt.startswith("i"):
            assert np.array_equal(c, want)
        else:
            assert rel_fro(c, want) <= (F32_TOL if dt == "f32" else F64_TOL)


def test_host_entry_k_pipelined(am, oracle):
    """Large row-major float32 products through am_host_gemm_strided_f32 take the K-pipelined schedule (K slices
    accumulated with beta = 1, then row chunks): uneven last row chunk, K not a multiple of the slice size, padded
    host rows; compared with the device-resident product and with oracle rows."""
    from arraymancer_b200 import _capi
    M, N, K = 4352, 4096, 4160
    a = rand((M, K + 8), "f32", 71)[:, :K]           # row pitch > K
    b = rand((K, N), "f32", 72)
    c = np.full((M, N + 16), np.float32(-5))          # row pitch > N: the gaps must survive
    cv = c[:, :N]
    lib = _capi.lib()
    _capi.check(lib.am_host_gemm_strided_f32(M, N, K, 1.0, a.ctypes.data, a.strides[0] // 4, 1, b.ctypes.data, N, 1, 0.0,
                                             cv.ctypes.data, c.strides[0] // 4, 1))
    assert np.all(c[:, N:] == -5)
    D = torch.empty((M, N), device="cuda")
    am.gemm_strided(1, dev(np.ascontiguousarray(a)), dev(b), 0, D)
    assert rel_fro(cv, D.cpu().numpy()) <= 2e-6
    rows = [0, 1, 543, 544, 2175, M - 1]
    want = oracle.matmul(np.ascontiguousarray(a[rows]), b)
    assert rel_fro(cv[rows], want) <= F32_TOL


def test_prepacked_operands(am, oracle):
    a, b = rand((700, 1000), "f32", 41), rand((1000, 600), "f32", 42)
    pa, pb = am.PackedF32(dev(a), "a"), am.PackedF32(dev(b), "b")
    R = oracle.matmul(a, b)
    for order in ("C", "F"):
        C = torch.empty((700, 600), device="cuda")
        if order == "F":
            C = C.t().contiguous().t()
        am.gemm_packed(1.0, pa, pb, 0.0, C)
        assert rel_fro(C.cpu().numpy(), R) <= F32_TOL
    a2 = rand((700, 1000), "f32", 43)
    pa.repack(dev(a2).t().contiguous().t())          # repack from a column-major view of new data
    C = torch.empty((700, 600), device="cuda")
    am.gemm_packed(1.0, pa, pb, 0.0, C)
    assert rel_fro(C.cpu().numpy(), oracle.matmul(a2, b)) <= F32_TOL
    pa.free(); pb.free()


def test_fused_allgather_epilogue_on_one_gpu(am, oracle):
    """am_gemm_packed_f32_bcast (row-sharded GEMM + all-gather in one kernel) with every "peer" copy of C on this GPU:
    exercises the local store + trickled peer copies of the epilogue, incl. edge tiles, several tiles per cluster,
    a short K (few chains per tile: the whole tile is forwarded at once) and a column-major C."""
    for (M, N, K) in [(700, 600, 1000), (2050, 1030, 96), (513, 4100, 2048)]:
        a, b = rand((M, K), "f32", 61), rand((K, N), "f32", 62)
        pa, pb = am.PackedF32(dev(a), "a"), am.PackedF32(dev(b), "b")
        R = oracle.matmul(a, b)
        for order in ("C", "F"):
            for npeers, me in ((1, 0), (3, 1), (8, 7)):
                # one buffer holding all copies, guard rows in between: nothing may be written outside the views
                full = torch.full((npeers, M + 2, N), -7.0, device="cuda")
                if order == "F":
                    full = full.permute(0, 2, 1).contiguous().permute(0, 2, 1)
                views = [full[g, 1:M + 1] for g in range(npeers)]
                am.gemm_packed_bcast(1.0, pa, pb, views[me], [v.data_ptr() for v in views], me)
                torch.cuda.synchronize()
                for g in range(npeers):
                    assert rel_fro(views[g].cpu().numpy(), R) <= F32_TOL, (M, N, K, order, npeers, g)
                    assert torch.equal(views[g], views[me])
                assert bool((full[:, 0] == -7.0).all()) and bool((full[:, M + 1] == -7.0).all())
        pa.free(); pb.free()


def test_cpp_host_mirror_known_answers():
    """The C++ host-side mirror (arraymancer_b200/host/arraymancer_b200.hpp) — CudaTensor, cuda(), `*`, gemm, conv2d,
    conv2d_backward over the C ABI — replays the reference's CUDA test vectors (tests/cpp/test_host_mirror.cpp)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "test_host_mirror")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/test_host_mirror not built (run __graft_entry__.build())")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.startswith("OK"), p.stdout + p.stderr
