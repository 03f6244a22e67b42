"""The `configs` block of bench.py: every BASELINE.json config measured in one driver run,
each record with its own timing (median of CUDA-event timed repetitions on the library's stream), roofline
fraction, nvidia-smi clock sample and a parity check against the CPU ORACLE (oracle/, test infrastructure — used
here only as the checker, outside every timed region).

Record layout:
  {"name", "config" (C1..C5 / "skinny"), "dtype", "shape", "ms" (median), "ms_best", "reps",
   "achieved", "peak", "unit", "frac", "bound" ("imad"|"dmma"|"tensor"|"hbm"|"ffma"), "peak_source",
   "gbps" (algorithmic bytes / ms), "l2": how cache effects are excluded, "clocks": {...},
   "parity": {"ok", "kind": "bit-exact"|"rel_fro", "value", "tol", "against", "sample"}}
Inputs follow SURVEY §8d: splitmix64 streams (seeds 42/43 C1, 7 full-range, 1234/1235 C3), the kostya generator for
f64, U[0,1) images / Kaiming-scaled weights for the LeNet convs.
"""
from __future__ import annotations

import time

import numpy as np
import torch

M64 = (1 << 64) - 1


# ---------------------------------------------------------------------------------------------- generators
def _splitmix64_torch(seed: int, n: int, device) -> torch.Tensor:
    """splitmix64 stream element i (1-based counter), computed on the device in wrapping int64 arithmetic."""
    def c(v):  # uint64 constant as a two's-complement python int
        return v - (1 << 64) if v >= (1 << 63) else v
    idx = torch.arange(1, n + 1, device=device, dtype=torch.int64)
    z = idx * c(0x9E3779B97F4A7C15) + c(seed & M64)
    z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * c(0xBF58476D1CE4E5B9)
    z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * c(0x94D049BB133111EB)
    return z ^ ((z >> 31) & ((1 << 33) - 1))


def splitmix64_numpy(seed: int, n: int) -> np.ndarray:
    """Host twin of _splitmix64_torch (used by the CPU tests to pin the device generator)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed & M64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def gen_matrix_chunked(kind, rows, cols, seed, device, dtype):
    """Row-major [rows, cols] device matrix, element (r, c) = f(splitmix64 stream `seed`, counter r*cols + c + 1),
    generated in chunks of 32 Mi elements.
    kind: 'u100' int U{0..99} (randomTensor(n,n,99), integer_matmul.nim:11) · 'full' full-range ints (wrap test) ·
          'u11' float U[-1,1) from the top 24 bits."""
    out = torch.empty((rows, cols), device=device, dtype=dtype)
    step = max(1, (1 << 25) // max(cols, 1))
    def c(v):
        return v - (1 << 64) if v >= (1 << 63) else v
    for r0 in range(0, rows, step):
        r1 = min(rows, r0 + step)
        idx = torch.arange(r0 * cols + 1, r1 * cols + 1, device=device, dtype=torch.int64)
        z = idx * c(0x9E3779B97F4A7C15) + c(seed & M64)
        z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * c(0xBF58476D1CE4E5B9)
        z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * c(0x94D049BB133111EB)
        z = z ^ ((z >> 31) & ((1 << 33) - 1))
        if kind == "u100":
            v = ((z >> 33) & ((1 << 31) - 1)) % 100
        elif kind == "full":
            v = z
        else:
            v = ((z >> 40) & ((1 << 24) - 1)).to(torch.float64) * (2.0 ** -23) - 1.0
        out[r0:r1] = v.to(dtype).reshape(r1 - r0, cols)
        del idx, z, v
    return out


def kostya(n: int, device, rows=None) -> torch.Tensor:
    """matgen of benchmarks/kostya_matmul.nim:6-11: a[i,j] = (1/n^2) * (i-j) * (i+j), float64."""
    i = (torch.arange(n, device=device, dtype=torch.float64) if rows is None
         else torch.as_tensor(rows, device=device, dtype=torch.float64)).reshape(-1, 1)
    j = torch.arange(n, device=device, dtype=torch.float64).reshape(1, -1)
    tmp = 1.0 / float(n * n)
    return (tmp * (i - j)) * (i + j)


# ---------------------------------------------------------------------------------------------- timing
class _Timer:
    def __init__(self, sampler_cls, gpu_index, flush_mb=256):
        self.sampler_cls, self.gpu_index = sampler_cls, gpu_index
        self.flush = torch.empty(flush_mb << 20, dtype=torch.uint8, device="cuda")

    def run(self, fn, flush_l2: bool, min_secs=0.7, min_reps=5, max_reps=200, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        sampler = self.sampler_cls(self.gpu_index)
        sampler.start()
        ts = []
        t_start = time.perf_counter()
        while len(ts) < max_reps and (len(ts) < min_reps or time.perf_counter() - t_start < min_secs):
            if flush_l2:
                self.flush.zero_()                       # 256 MB > 126 MB L2: evicts the operands
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        clocks = sampler.stop()
        ts.sort()
        return ts[len(ts) // 2], ts[0], len(ts), clocks


def _rel(got: np.ndarray, want: np.ndarray) -> float:
    g, w = got.astype(np.float64), want.astype(np.float64)
    return float(np.linalg.norm(g - w) / max(np.linalg.norm(w), 1e-300))


def _rec(name, config, dtype, shape, ms, best, reps, clocks, ops, bytes_, peak, unit_scale, unit, bound, peak_source,
         parity, l2, extra=None):
    achieved = ops / (ms * 1e-3) / unit_scale if bound != "hbm" else bytes_ / (ms * 1e-3) / 1e9
    r = {"name": name, "config": config, "dtype": dtype, "shape": shape, "ms": ms, "ms_best": best, "reps": reps,
         "achieved": achieved, "peak": peak, "unit": unit if bound != "hbm" else "GB/s",
         "frac": (achieved / peak) if peak else None, "bound": bound, "peak_source": peak_source,
         "ops": ops, "algorithmic_bytes": bytes_, "gbps": bytes_ / (ms * 1e-3) / 1e9,
         "l2": l2, "clocks": clocks, "parity": parity}
    if bound != "hbm":
        r["top_s"] = ops / (ms * 1e-3) / 1e12
    if extra:
        r.update(extra)
    return r


# ---------------------------------------------------------------------------------------------- the block
def run_configs(am, orc, peaks, sampler_cls, gpu_index=0, only=None, log=None):
    """Returns (records, pipe_peaks).  `peaks` = MEASURED_PEAKS.json dict (hbm_gbs, bf16_tflops...)."""
    from arraymancer_b200 import _capi
    dev = torch.device("cuda", torch.cuda.current_device())
    say = log or (lambda *_: None)
    recs = []
    T = _Timer(sampler_cls, gpu_index)
    FL = "L2 flushed (256 MB memset) before every timed repetition"
    BIG = "operands >> 126 MB L2, no flush needed"

    # ---- per-dtype pipe peaks from the library's own micro-benchmarks (measured on the box, in this run)
    pipe = {}
    for i, n in enumerate(["ffma_f32", "dfma_f64", "imad_i32", "i64_mac", "dmma_f64", "umma_tf32_1cta", "umma_tf32_2cta",
                           "i64_narrow_mac"]):
        try:
            pipe[n] = _capi.microbench(i)
        except Exception as e:  # noqa: BLE001
            pipe[n] = None
            say(f"microbench {n} failed: {e}")
    for n, i in (("ffma2_f32x2", 13), ("umma_tf32_n64_conv_issue_pattern", 31), ("umma_tf32_n64_conv_pattern_with_weight_stream", 32)):
        try:
            pipe[n] = _capi.microbench(i)
        except Exception as e:  # noqa: BLE001
            pipe[n] = None
            say(f"microbench {n} failed: {e}")
    hbm = float(peaks.get("hbm_gbs", 6454.3))
    tf32x3 = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0))) / 6.0
    P = {"f32": (tf32x3, "TFLOP/s", "tensor", "MEASURED_PEAKS bf16_tflops_sustained / 2 / 3 (3xTF32)"),
         "f64": (pipe.get("dmma_f64") or 37.0, "TFLOP/s", "dmma", "am_microbench 4 (DMMA m8n8k4), this run"),
         "i32": (pipe.get("imad_i32") or 36.1, "TOP/s", "imad", "am_microbench 2 (IMAD), this run"),
         "i64": (pipe.get("i64_mac") or 7.57, "TOP/s", "imad", "am_microbench 3 (IMAD.WIDE+2 IMAD per mac), this run"),
         "i64n": (pipe.get("i64_narrow_mac") or 15.3, "TOP/s", "imad",
                  "am_microbench 7 (one IMAD.WIDE per mac, int32-range operands), this run")}
    tdt = {"f32": torch.float32, "f64": torch.float64, "i32": torch.int32, "i64": torch.int64}

    def want(tag):
        return only is None or any(tag.startswith(o) for o in only)

    def gemm_record(name, config, dt, M, N, K, A, B, C, peak_key, flush, parity_rows, tol=None, alpha=1, beta=0,
                    note=None):
        """A, B, C: device views (any strides).  parity_rows: row indices compared with the oracle (None = all)."""
        ms, best, reps, clocks = T.run(lambda: am.gemm_strided(alpha, A, B, beta, C), flush)
        peak, unit, bound, src = P[peak_key]
        # ---- parity vs the oracle (outside the timed region), on the very C the timed runs produced
        rows = np.arange(M) if parity_rows is None else np.asarray(parity_rows)
        ridx = torch.as_tensor(rows, device=dev)
        a_h = A.index_select(0, ridx).cpu().numpy()
        b_h = B.cpu().numpy()
        got = C.index_select(0, ridx).cpu().numpy()
        wantc = np.zeros((len(rows), N), dtype=a_h.dtype)
        orc.gemm_strided(alpha, np.ascontiguousarray(a_h), b_h, 0, wantc)
        if dt in ("i32", "i64"):
            ok = bool(np.array_equal(got, wantc))
            parity = {"ok": ok, "kind": "bit-exact", "value": int((got != wantc).sum()), "tol": 0}
        else:
            rel = _rel(got, wantc)
            parity = {"ok": bool(rel <= tol and np.isfinite(got).all()), "kind": "rel_fro", "value": rel, "tol": tol}
        parity["against"] = "oracle.gemm_strided (restated laser gemm_strided, default ISA variant)"
        parity["sample"] = "all rows" if parity_rows is None else f"{len(rows)} rows of C (full N, full K)"
        sz = A.element_size()
        r = _rec(name, config, dt, [M, N, K], ms, best, reps, clocks, 2.0 * M * N * K, sz * (M * K + K * N + M * N), peak,
                 1e12, unit, bound, src, parity, FL if flush else BIG, {"note": note} if note else None)
        recs.append(r)
        say(f"{name}: {ms:.3f} ms  {r['achieved']:.2f} {r['unit']}  frac {r['frac']:.3f}  parity {parity['ok']} ({parity['value']})")
        del a_h, b_h, got, wantc
        return r

    def sample_rows(M, k=64, seed=5):
        rng = np.random.default_rng(seed)
        return np.sort(rng.choice(M, size=min(k, M), replace=False))

    # ================================================================ C1: int64 1500^2 (+ 8192^2, int32)
    if want("C1"):
        for n, narrow in ((1500, True), (1500, False), (8192, True), (8192, False)):
            kind = "u100" if narrow else "full"
            sa, sb = (42, 43) if narrow else (7, 8)
            A = gen_matrix_chunked(kind, n, n, sa, dev, torch.int64)
            B = gen_matrix_chunked(kind, n, n, sb, dev, torch.int64)
            C = torch.empty((n, n), device=dev, dtype=torch.int64)
            gemm_record(f"C1_i64_{n}_{'u100' if narrow else 'fullrange'}", "C1", "i64", n, n, n, A, B, C,
                        "i64n" if narrow else "i64", flush=(n <= 2048), parity_rows=None if n <= 2048 else sample_rows(n),
                        note=("benchmarks/integer_matmul.nim:11-12 inputs U{0..99}: int32-range operands take the "
                              "one-IMAD.WIDE path" if narrow else "full-range operands: wrap mod 2^64"))
            del A, B, C
        for n in (1500, 8192):
            A = gen_matrix_chunked("full", n, n, 7, dev, torch.int32)
            B = gen_matrix_chunked("full", n, n, 8, dev, torch.int32)
            C = torch.empty((n, n), device=dev, dtype=torch.int32)
            gemm_record(f"C1_i32_{n}_fullrange", "C1", "i32", n, n, n, A, B, C, "i32", flush=(n <= 2048),
                        parity_rows=None if n <= 2048 else sample_rows(n), note="wrap mod 2^32")
            del A, B, C
        torch.cuda.empty_cache()

    # ================================================================ C2: float64 kostya, 8192^2 (and 1500^2)
    if want("C2"):
        for n in (1500, 8192):
            A = kostya(n, dev)
            B = A.clone()
            C = torch.empty((n, n), device=dev, dtype=torch.float64)
            gemm_record(f"C2_f64_{n}_kostya", "C2", "f64", n, n, n, A, B, C, "f64", flush=(n <= 2048),
                        parity_rows=None if n <= 2048 else sample_rows(n), tol=1e-13,
                        note="benchmarks/kostya_matmul.nim:6-19 matgen")
            del A, B, C
        torch.cuda.empty_cache()

    # ================================================================ C3: float32 16384^2 and its strided / transposed views
    if want("C3"):
        n = 16384
        A = gen_matrix_chunked("u11", n, n, 1234, dev, torch.float32)
        B = gen_matrix_chunked("u11", n, n, 1235, dev, torch.float32)
        C = torch.empty((n, n), device=dev, dtype=torch.float32)
        rows = sample_rows(n)
        gemm_record("C3_f32_16384_rowmajor", "C3", "f32", n, n, n, A, B, C, "f32", False, rows, tol=5e-6)
        At = A.t()            # logical A^T: (rs=1, cs=n) view of the same buffer
        gemm_record("C3_f32_16384_At_view", "C3", "f32", n, n, n, At, B, C, "f32", False, rows, tol=5e-6,
                    note="A is a transposed view (rowStride 1, colStride n)")
        Bt = B.t()
        gemm_record("C3_f32_16384_Bt_view", "C3", "f32", n, n, n, A, Bt, C, "f32", False, rows, tol=5e-6,
                    note="B is a transposed view")
        gemm_record("C3_f32_16384_At_Bt_views", "C3", "f32", n, n, n, At, Bt, C, "f32", False, rows, tol=5e-6,
                    note="both operands transposed views")
        Cc = torch.empty((n, n), device=dev, dtype=torch.float32).t()   # column-major C (CudaTensor default layout)
        gemm_record("C3_f32_16384_C_colmajor", "C3", "f32", n, n, n, A, B, Cc, "f32", False, rows, tol=5e-6,
                    note="C column-major (rowStride 1, colStride n): the CudaTensor default")
        del At, Bt, Cc, A
        torch.cuda.empty_cache()
        parent = gen_matrix_chunked("u11", 2 * n, n, 1236, dev, torch.float32)
        A2 = parent[::2]      # step-2 row slice of a 2n x n parent: rowStride 2n
        gemm_record("C3_f32_16384_step2_rows", "C3", "f32", n, n, n, A2, B, C, "f32", False, rows, tol=5e-6,
                    note="A = every second row of a 2n x n parent (rowStride 2n, colStride 1)")
        del parent, A2, B, C
        torch.cuda.empty_cache()

    # ================================================================ skinny / HBM-bound GEMMs (DRAM GB/s evidence)
    if want("skinny"):
        n = 16384
        B = gen_matrix_chunked("u11", n, n, 1235, dev, torch.float32)
        for m in (1, 16, 64):
            A = gen_matrix_chunked("u11", m, n, 1234, dev, torch.float32)
            C = torch.empty((m, n), device=dev, dtype=torch.float32)
            ms, best, reps, clocks = T.run(lambda: am.gemm_strided(1, A, B, 0, C), False)
            got = C.cpu().numpy()
            wantc = np.zeros_like(got)
            orc.gemm_strided(1, A.cpu().numpy(), B.cpu().numpy(), 0, wantc)
            rel = _rel(got, wantc)
            bytes_ = 4 * (m * n + n * n + m * n)
            # thin dimension <= 16: 2*m flop per 4-byte element of B is below the FFMA ridge -> DRAM-bound, reported in GB/s;
            # m = 64: 128 flop per element, FFMA-bound on the exact-fp32 SIMT kernel -> reported against the FFMA pipe
            ffma_bound = m > 16 and pipe.get("ffma_f32")
            recs.append(_rec(f"skinny_f32_M{m}_N{n}_K{n}", "skinny", "f32", [m, n, n], ms, best, reps, clocks, 2.0 * m * n * n,
                             bytes_, pipe["ffma_f32"] if ffma_bound else hbm, 1e12, "TFLOP/s" if ffma_bound else "GB/s",
                             "ffma" if ffma_bound else "hbm", "am_microbench 0 (FFMA), this run" if ffma_bound else "MEASURED_PEAKS hbm_gbs",
                             {"ok": bool(rel <= 5e-6), "kind": "rel_fro", "value": rel, "tol": 5e-6,
                              "against": "oracle.gemm_strided", "sample": "all rows"}, BIG,
                             {"note": "M rows against a 1 GiB B: the product streams B once (DRAM-bound)"}))
            say(f"skinny M={m}: {ms:.3f} ms {recs[-1]['achieved']:.0f} {recs[-1]['unit']} frac {recs[-1]['frac']:.3f} rel {rel:.2e}")
            # the transposed twin: tall A (n x n) times a skinny B (n x m), i.e. the gemv-like `A * v` of the reference
            Bs = gen_matrix_chunked("u11", n, m, 1237, dev, torch.float32)
            Cs = torch.empty((n, m), device=dev, dtype=torch.float32)
            ms, best, reps, clocks = T.run(lambda: am.gemm_strided(1, B, Bs, 0, Cs), False)
            got = Cs.cpu().numpy()
            wantc = np.zeros_like(got)
            orc.gemm_strided(1, B.cpu().numpy(), Bs.cpu().numpy(), 0, wantc)
            rel = _rel(got, wantc)
            recs.append(_rec(f"skinny_f32_M{n}_N{m}_K{n}", "skinny", "f32", [n, m, n], ms, best, reps, clocks, 2.0 * m * n * n,
                             bytes_, pipe["ffma_f32"] if ffma_bound else hbm, 1e12, "TFLOP/s" if ffma_bound else "GB/s",
                             "ffma" if ffma_bound else "hbm", "am_microbench 0 (FFMA), this run" if ffma_bound else "MEASURED_PEAKS hbm_gbs",
                             {"ok": bool(rel <= 5e-6), "kind": "rel_fro", "value": rel, "tol": 5e-6,
                              "against": "oracle.gemm_strided", "sample": "all rows"}, BIG,
                             {"note": "1 GiB A against N columns (N = 1: the matrix-vector product of `*`)"}))
            say(f"skinny N={m}: {ms:.3f} ms {recs[-1]['achieved']:.0f} {recs[-1]['unit']} frac {recs[-1]['frac']:.3f} rel {rel:.2e}")
            del A, C, Bs, Cs
        del B
        torch.cuda.empty_cache()

    # ================================================================ C4: LeNet conv2d fwd + bwd, batch 4096
    if want("C4"):
        NB = 4096
        layers = [("cv1", (NB, 1, 28, 28), (20, 1, 5, 5), 2024), ("cv2", (NB, 20, 12, 12), (50, 20, 5, 5), 2025)]
        for lname, xs, ks, seed in layers:
            g = torch.Generator(device=dev); g.manual_seed(seed)
            X = torch.rand(xs, device=dev, dtype=torch.float32, generator=g)
            fan_in = ks[1] * ks[2] * ks[3]
            W = torch.randn(ks, device=dev, dtype=torch.float32, generator=g) * float(np.sqrt(2.0 / fan_in))
            Bv = torch.rand((ks[0], 1, 1), device=dev, dtype=torch.float32, generator=g) * 0.1
            out = am.conv2d(X, W, Bv)
            GO = torch.rand(out.shape, device=dev, dtype=torch.float32, generator=g) * 2 - 1
            Ho, Wo = out.shape[2], out.shape[3]
            flops = 2.0 * NB * ks[0] * Ho * Wo * fan_in
            by_f = 4 * (X.numel() + W.numel() + ks[0] + out.numel())
            by_b = 4 * (X.numel() + W.numel() + GO.numel() + X.numel() + W.numel() + ks[0])
            x_h, w_h, b_h, go_h = X.cpu().numpy(), W.cpu().numpy(), Bv.cpu().numpy(), GO.cpu().numpy()
            # forward
            ms, best, reps, clocks = T.run(lambda: am.conv2d(X, W, Bv), flush_l2=True)
            t0 = time.perf_counter()
            want_f = orc.conv2d(x_h, w_h, b_h)
            rel = _rel(am.conv2d(X, W, Bv).cpu().numpy(), want_f)
            par = {"ok": bool(rel <= 5e-6), "kind": "rel_fro", "value": rel, "tol": 5e-6,
                   "against": "oracle.conv2d (restated im2colgemm_conv2d)", "sample": f"all {NB} images",
                   "oracle_seconds": time.perf_counter() - t0}
            hbm_bound = (flops / by_f) < (tf32x3 * 1e12 / (hbm * 1e9))
            if hbm_bound:
                recs.append(_rec(f"C4_{lname}_fwd", "C4", "f32", list(xs) + list(ks), ms, best, reps, clocks, flops, by_f, hbm,
                                 1e12, "GB/s", "hbm", "MEASURED_PEAKS hbm_gbs", par, FL))
            else:
                recs.append(_rec(f"C4_{lname}_fwd", "C4", "f32", list(xs) + list(ks), ms, best, reps, clocks, flops, by_f,
                                 tf32x3, 1e12, "TFLOP/s", "tensor", P["f32"][3], par, FL))
            say(f"{lname} fwd: {ms:.4f} ms frac {recs[-1]['frac']:.3f} rel {rel:.2e}")
            # backward (data + weight + bias gradients in one call)
            ms, best, reps, clocks = T.run(lambda: am.conv2d_backward(X, W, Bv, (0, 0), (1, 1), (1, 1), GO), flush_l2=True)
            gi, gw, gb = am.conv2d_backward(X, W, Bv, (0, 0), (1, 1), (1, 1), GO)
            t0 = time.perf_counter()
            wgi, wgw, wgb = orc.conv2d_backward(x_h, w_h, go_h)
            rels = {"grad_input": _rel(gi.cpu().numpy(), wgi), "grad_kernel": _rel(gw.cpu().numpy(), wgw),
                    "grad_bias": _rel(gb.cpu().numpy(), wgb)}
            par = {"ok": bool(rels["grad_input"] <= 5e-6 and rels["grad_kernel"] <= 1e-4 and rels["grad_bias"] <= 1e-4),
                   "kind": "rel_fro", "value": max(rels.values()), "values": rels,
                   "tol": {"grad_input": 5e-6, "grad_kernel": 1e-4, "grad_bias": 1e-4},
                   "against": "oracle.conv2d_backward (restated im2colgemm_conv2d_gradient + grad_bias)",
                   "sample": f"all {NB} images", "oracle_seconds": time.perf_counter() - t0}
            fl_b = 2 * flops + NB * ks[0] * Ho * Wo
            hbm_bound = (fl_b / by_b) < (tf32x3 * 1e12 / (hbm * 1e9))
            if hbm_bound:
                recs.append(_rec(f"C4_{lname}_bwd", "C4", "f32", list(xs) + list(ks), ms, best, reps, clocks, fl_b, by_b, hbm,
                                 1e12, "GB/s", "hbm", "MEASURED_PEAKS hbm_gbs", par, FL))
            else:
                recs.append(_rec(f"C4_{lname}_bwd", "C4", "f32", list(xs) + list(ks), ms, best, reps, clocks, fl_b, by_b,
                                 tf32x3, 1e12, "TFLOP/s", "tensor", P["f32"][3], par, FL))
            say(f"{lname} bwd: {ms:.4f} ms frac {recs[-1]['frac']:.3f} rels {rels}")
            del X, W, Bv, out, GO, gi, gw, gb
        torch.cuda.empty_cache()

    # ================================================================ C5: float64 32768^2 (the f32 twin is bench.py's headline)
    if want("C5"):
        n = 32768
        A = kostya(n, dev)
        B = A.clone()
        C = torch.empty((n, n), device=dev, dtype=torch.float64)
        ms, best, reps, clocks = T.run(lambda: am.gemm_strided(1.0, A, B, 0.0, C), False, min_secs=0.0, min_reps=3, warm=1)
        rows = sample_rows(n, 32)
        ridx = torch.as_tensor(rows, device=dev)
        a_h = A.index_select(0, ridx).cpu().numpy()
        got = C.index_select(0, ridx).cpu().numpy()
        del A, C
        b_h = B.cpu().numpy()
        del B
        torch.cuda.empty_cache()
        wantc = np.zeros_like(got)
        orc.gemm_strided(1.0, a_h, b_h, 0.0, wantc)
        rel = _rel(got, wantc)
        peak, unit, bound, src = P["f64"]
        recs.append(_rec("C5_f64_32768_kostya", "C5", "f64", [n, n, n], ms, best, reps, clocks, 2.0 * n ** 3, 8 * 3 * n * n, peak,
                         1e12, unit, bound, src,
                         {"ok": bool(rel <= 1e-13), "kind": "rel_fro", "value": rel, "tol": 1e-13,
                          "against": "oracle.gemm_strided", "sample": "32 rows of C (full N, full K = 32768)"}, BIG))
        say(f"C5 f64 32768: {ms:.1f} ms {recs[-1]['achieved']:.2f} TFLOP/s frac {recs[-1]['frac']:.3f} rel {rel:.2e}")
        del a_h, b_h, got, wantc
    return recs, pipe
