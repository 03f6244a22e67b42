"""GPU bring-up / measurement driver: runs each experiment in its own subprocess with a timeout so a
hung kernel cannot take the whole gpurun call down.  Writes JSON lines to gpurun_out/bringup.jsonl.

  python tools/gpu_bringup.py            # everything
  python tools/gpu_bringup.py one NAME   # one experiment in-process (used by the parent)
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def _time_gpu(fn, warm=2, reps=5):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def exp_peaks():
    from arraymancer_b200 import _capi
    names = ["ffma_f32", "dfma_f64", "imad_i32", "i64_mac", "dmma_f64", "umma_tf32_1cta", "umma_tf32_2cta", "i64_narrow_mac"]
    res = {}
    for i, n in enumerate(names):
        try:
            res[n] = _capi.microbench(i)
        except Exception as e:  # noqa
            res[n] = f"ERR {e}"
    return res


def exp_peaks_small_n():
    from arraymancer_b200 import _capi
    names = {8: "umma_n64", 9: "umma_n64_2acc", 10: "umma_n64_ts", 11: "umma_n64_ts_2acc", 12: "umma_n128"}
    return {n: _capi.microbench(i) for i, n in names.items()}


def _rand(shape, dt, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    if dt in ("f32", "f64"):
        return (rng.random(shape) * 2 - 1).astype({"f32": np.float32, "f64": np.float64}[dt])
    if dt == "i32":
        return rng.integers(-2**31, 2**31 - 1, size=shape, dtype=np.int64).astype(np.int32)
    return rng.integers(-2**63, 2**63 - 1, size=shape, dtype=np.int64)


def _parity(dt, M, N, K, path=None, alpha=1, beta=0, layout="rr"):
    import numpy as np
    import torch
    import arraymancer_b200 as am
    from oracle import laser_oracle as orc
    if path is not None:
        am.set_f32_path(path)
    a, b = _rand((M, K), dt, 1), _rand((K, N), dt, 2)
    c0 = _rand((M, N), dt, 3)
    A = torch.from_numpy(a).cuda()
    B = torch.from_numpy(b).cuda()
    if layout[0] == "c":
        A = A.t().contiguous().t()
    if layout[1] == "c":
        B = B.t().contiguous().t()
    C = torch.from_numpy(c0).cuda()
    if len(layout) > 2 and layout[2] == "c":
        C = C.t().contiguous().t()
    am.gemm_strided(alpha, A, B, beta, C)
    torch.cuda.synchronize()
    got = C.cpu().numpy()
    want = c0.copy()
    orc.gemm_strided(alpha, a, b, beta, want)
    if dt in ("i32", "i64"):
        return {"exact": bool(np.array_equal(got, want)), "mismatch": int((got != want).sum())}
    rel = float(np.linalg.norm(got.astype(np.float64) - want.astype(np.float64)) / np.linalg.norm(want.astype(np.float64)))
    return {"rel_fro": rel, "finite": bool(np.isfinite(got).all())}


def exp_simt_parity():
    res = {}
    for dt in ("i64", "i32", "f64", "f32"):
        for (M, N, K) in [(5, 7, 3), (64, 64, 64), (129, 65, 300), (257, 513, 100), (1500, 1500, 1500)]:
            for lay in ("rr", "cc", "rrc"):
                res[f"{dt}_{M}x{N}x{K}_{lay}"] = _parity(dt, M, N, K, path=1, alpha=1, beta=0, layout=lay)
        res[f"{dt}_alpha_beta"] = _parity(dt, 100, 90, 80, path=1, alpha=-3, beta=2)
    return res


def _tc_parity(path, flush):
    os.environ["AM_TC_FLUSH_KB"] = str(flush)
    res = {}
    for (M, N, K) in [(128, 256, 32), (256, 512, 64), (300, 700, 100), (1024, 2048, 1024), (512, 768, 4096), (512, 512, 16384)]:
        for lay in ("rr", "rrc", "cc"):
            res[f"{M}x{N}x{K}_{lay}"] = _parity("f32", M, N, K, path=path, layout=lay)
    res["alpha_beta"] = _parity("f32", 512, 512, 512, path=path, alpha=-3, beta=2)
    return res


def exp_tc1_f2(): return _tc_parity(3, 2)
def exp_tc2_f1(): return _tc_parity(2, 1)
def exp_tc2_f2(): return _tc_parity(2, 2)
def exp_tc2_f4(): return _tc_parity(2, 4)
def exp_tc2_f8(): return _tc_parity(2, 8)


def _gemm_speed(dt, n, path=None, reps=5, wide=False):
    import torch
    import arraymancer_b200 as am
    if path is not None:
        am.set_f32_path(path)
    tdt = {"f32": torch.float32, "f64": torch.float64, "i32": torch.int32, "i64": torch.int64}[dt]
    if dt in ("f32", "f64"):
        A = torch.rand(n, n, device="cuda", dtype=tdt) * 2 - 1
        B = torch.rand(n, n, device="cuda", dtype=tdt) * 2 - 1
    elif wide:
        A = torch.randint(-2**62, 2**62, (n, n), device="cuda", dtype=tdt)
        B = torch.randint(-2**62, 2**62, (n, n), device="cuda", dtype=tdt)
    else:
        A = torch.randint(0, 100, (n, n), device="cuda", dtype=tdt)
        B = torch.randint(0, 100, (n, n), device="cuda", dtype=tdt)
    C = torch.empty(n, n, device="cuda", dtype=tdt)
    med, best = _time_gpu(lambda: am.gemm_strided(1, A, B, 0, C), reps=reps)
    return {"ms_med": med, "ms_best": best, "tops_med": 2 * n**3 / med / 1e9, "tops_best": 2 * n**3 / best / 1e9}


def exp_simt_speed():
    res = {}
    for dt, n in [("i64", 1500), ("i64", 4096), ("i32", 1500), ("i32", 4096), ("f64", 1500), ("f64", 4096),
                  ("f64", 8192), ("f32", 4096)]:
        res[f"{dt}_{n}"] = _gemm_speed(dt, n, path=1, reps=3)
    for n in (1500, 4096, 8192):
        res[f"i64_fullrange_{n}"] = _gemm_speed("i64", n, reps=3, wide=True)
    res["i64_8192"] = _gemm_speed("i64", 8192, reps=3)
    res["i32_8192"] = _gemm_speed("i32", 8192, reps=3)
    return res


def _tc_speed(path, flush):
    os.environ["AM_TC_FLUSH_KB"] = str(flush)
    res = {}
    for n in (4096, 8192, 16384):
        res[f"f32_{n}"] = _gemm_speed("f32", n, path=path, reps=3)
    return res


def exp_tc1_speed_f2(): return _tc_speed(3, 2)
def exp_tc2_speed_f1(): return _tc_speed(2, 1)
def exp_tc2_speed_f2(): return _tc_speed(2, 2)
def exp_tc2_speed_f4(): return _tc_speed(2, 4)
def exp_tc2_speed_f8(): return _tc_speed(2, 8)


def exp_f64_speed():
    import arraymancer_b200 as am
    res = {}
    for name, path in (("dmma", am.F64_DMMA), ("simt", am.F64_SIMT)):
        am.set_f64_path(path)
        for n in (1500, 4096, 8192):
            res[f"{name}_{n}"] = _gemm_speed("f64", n, reps=3)
    am.set_f64_path(am.F64_AUTO)
    return res


def exp_conv_sweep():
    """tile-plan sweep of the direct conv kernels (env overrides read at plan time)"""
    import torch
    import arraymancer_b200 as am
    res = {}
    layers = [("cv1", (4096, 1, 28, 28), (20, 1, 5, 5)), ("cv2", (4096, 20, 12, 12), (50, 20, 5, 5))]
    for name, xs, ks in layers:
        X = torch.rand(xs, device="cuda"); K_ = torch.randn(ks, device="cuda") * 0.1; B_ = torch.zeros(ks[0], 1, 1, device="cuda")
        out = am.conv2d(X, K_, B_); go = torch.ones_like(out)
        for maxt in (128, 256, 384, 512):
            for ct, px in ((4, 4), (4, 8), (8, 4), (8, 8)):
                for kb in (24, 40, 80):
                    os.environ.update(AM_CONV_MAXT=str(maxt), AM_CONV_CT=str(ct), AM_CONV_PX=str(px), AM_CONV_SMEMKB=str(kb))
                    try:
                        f, _ = _time_gpu(lambda: am.conv2d(X, K_, B_), warm=1, reps=3)
                        d, _ = _time_gpu(lambda: am.conv2d_backward(X, K_, B_, (0, 0), (1, 1), (1, 1), go, need_kernel_grad=False), warm=1, reps=3)
                        res[f"{name}_t{maxt}_ct{ct}_px{px}_kb{kb}"] = [round(f, 4), round(d, 4)]
                    except Exception as e:  # noqa
                        res[f"{name}_t{maxt}_ct{ct}_px{px}_kb{kb}"] = str(e)[:60]
    return res


def exp_conv_parity():
    import numpy as np
    import torch
    import arraymancer_b200 as am
    from oracle import laser_oracle as orc
    res = {}
    rng = np.random.default_rng(0)
    cases = [((2, 3, 4, 5), (2, 3, 3, 3), (1, 1), (1, 1), (1, 1)), ((3, 1, 28, 28), (20, 1, 5, 5), (0, 0), (1, 1), (1, 1)),
             ((2, 20, 12, 12), (50, 20, 5, 5), (0, 0), (1, 1), (1, 1)), ((2, 4, 9, 8), (5, 4, 3, 2), (1, 2), (2, 1), (1, 1)),
             ((2, 3, 11, 10), (70, 3, 3, 3), (2, 2), (1, 1), (2, 2)), ((1, 130, 6, 6), (6, 130, 3, 3), (1, 0), (2, 2), (2, 1))]
    for dt in ("f32", "f64", "i32", "i64"):
        for ci, (xs, ks, pad, st, dil) in enumerate(cases):
            if dt in ("f32", "f64"):
                npdt = np.float32 if dt == "f32" else np.float64
                x = rng.random(xs).astype(npdt); k = (rng.random(ks) - 0.5).astype(npdt); b = rng.random((ks[0], 1, 1)).astype(npdt)
            else:
                npdt = np.int32 if dt == "i32" else np.int64
                x = rng.integers(-9, 9, xs).astype(npdt); k = rng.integers(-9, 9, ks).astype(npdt); b = rng.integers(-9, 9, (ks[0], 1, 1)).astype(npdt)
            want = orc.conv2d(x, k, b, pad, st, dil)
            X, K_, B_ = (torch.from_numpy(v).cuda() for v in (x, k, b))
            got = am.conv2d(X, K_, B_, pad, st, dil).cpu().numpy()
            go = (rng.random(want.shape) * 2 - 1).astype(npdt) if dt in ("f32", "f64") else rng.integers(-5, 5, want.shape).astype(npdt)
            wgi, wgw, wgb = orc.conv2d_backward(x, k, go, True, pad, st, dil)
            gi, gw, gb = am.conv2d_backward(X, K_, B_, pad, st, dil, torch.from_numpy(go).cuda())
            gi, gw, gb = gi.cpu().numpy(), gw.cpu().numpy(), gb.cpu().numpy()
            if dt in ("i32", "i64"):
                res[f"{dt}_{ci}"] = {"fwd": bool(np.array_equal(got, want)), "gin": bool(np.array_equal(gi, wgi)),
                                     "gw": bool(np.array_equal(gw, wgw)), "gb": bool(np.array_equal(gb, wgb))}
            else:
                f = lambda a_, b_: float(np.linalg.norm(a_.astype(np.float64) - b_) / max(np.linalg.norm(b_.astype(np.float64)), 1e-30))
                res[f"{dt}_{ci}"] = {"fwd": f(got, want), "gin": f(gi, wgi), "gw": f(gw, wgw), "gb": f(gb, wgb)}
    return res


def exp_conv_speed():
    import torch
    import arraymancer_b200 as am
    res = {}
    if os.environ.get("AM_BRINGUP_CONV_PATH"):
        am._capi.set_conv_path(int(os.environ["AM_BRINGUP_CONV_PATH"]))
    for name, xs, ks in [("cv1", (4096, 1, 28, 28), (20, 1, 5, 5)), ("cv2", (4096, 20, 12, 12), (50, 20, 5, 5))]:
        X = torch.rand(xs, device="cuda"); K_ = torch.randn(ks, device="cuda") * 0.1; B_ = torch.zeros(ks[0], 1, 1, device="cuda")
        out = am.conv2d(X, K_, B_)
        go = torch.ones_like(out)
        med, best = _time_gpu(lambda: am.conv2d(X, K_, B_), reps=5)
        bmed, bbest = _time_gpu(lambda: am.conv2d_backward(X, K_, B_, (0, 0), (1, 1), (1, 1), go), reps=5)
        dmed, _ = _time_gpu(lambda: am.conv2d_backward(X, K_, B_, (0, 0), (1, 1), (1, 1), go, need_kernel_grad=False), reps=5)
        wmed, _ = _time_gpu(lambda: am.conv2d_backward(X, K_, B_, (0, 0), (1, 1), (1, 1), go, need_input_grad=False), reps=5)
        flops = 2 * out.numel() * ks[1] * ks[2] * ks[3]
        res[name] = {"fwd_ms": med, "fwd_gflops": flops / med / 1e6, "bwd_ms": bmed, "bwd_gflops": 2 * flops / bmed / 1e6,
                     "dgrad_ms": dmed, "wgrad_ms": wmed}
    return res


def exp_nn_speed():
    """HBM-bound LeNet companions (batch 4096): achieved GB/s (algorithmic bytes) and a resident LeNet fwd+bwd step."""
    import torch
    import arraymancer_b200 as am
    res = {}
    B = 4096
    x = torch.rand((B, 20, 24, 24), device="cuda") - 0.5
    g = torch.rand_like(x)
    nb = x.numel() * 4
    med, _ = _time_gpu(lambda: am.relu(x), reps=5); res["relu_fwd_cv1"] = {"ms": med, "GBps": 2 * nb / med / 1e6}
    med, _ = _time_gpu(lambda: am.relu_backward(g, x), reps=5); res["relu_bwd_cv1"] = {"ms": med, "GBps": 3 * nb / med / 1e6}
    idx, p = am.maxpool2d(x, (2, 2), (0, 0), (2, 2))
    med, _ = _time_gpu(lambda: am.maxpool2d(x, (2, 2), (0, 0), (2, 2)), reps=5)
    res["maxpool_fwd_cv1"] = {"ms": med, "GBps": (nb + p.numel() * 4 + idx.numel() * 8) / med / 1e6}
    gp = torch.rand_like(p)
    med, _ = _time_gpu(lambda: am.maxpool2d_backward(x.shape, idx, gp, windows_overlap=False), reps=5)
    res["maxpool_bwd_cv1"] = {"ms": med, "GBps": (nb + p.numel() * 4 + idx.numel() * 8) / med / 1e6}
    f = torch.rand((B, 800), device="cuda"); w = torch.rand((500, 800), device="cuda") - 0.5; b = torch.rand((1, 500), device="cuda")
    med, _ = _time_gpu(lambda: am.linear(f, w, b), reps=5); res["linear_fwd_800_500"] = {"ms": med, "tflops": 2 * B * 800 * 500 / med / 1e9}
    go = torch.rand((B, 500), device="cuda")
    med, _ = _time_gpu(lambda: am.linear_backward(f, w, go), reps=5); res["linear_bwd_800_500"] = {"ms": med, "tflops": 4 * B * 800 * 500 / med / 1e9}
    am.set_f32_path(am.F32_TC)
    try:
        med, _ = _time_gpu(lambda: am.linear(f, w, b), reps=5); res["linear_fwd_800_500_tc"] = {"ms": med, "tflops": 2 * B * 800 * 500 / med / 1e9}
        med, _ = _time_gpu(lambda: am.linear_backward(f, w, go), reps=5); res["linear_bwd_800_500_tc"] = {"ms": med, "tflops": 4 * B * 800 * 500 / med / 1e9}
    finally:
        am.set_f32_path(am.F32_AUTO)
    lg = torch.rand((B, 10), device="cuda"); lab = torch.randint(0, 10, (B,), device="cuda")
    med, _ = _time_gpu(lambda: am.sparse_softmax_cross_entropy_dev(lg, lab), reps=5); res["ssce_fwd"] = {"ms": med}
    med, _ = _time_gpu(lambda: am.sparse_softmax_cross_entropy_backward(1.0, lg, lab), reps=5); res["ssce_bwd"] = {"ms": med}

    # resident LeNet step (ex02_mnist.nim network), forward + backward, batch 4096
    X = torch.rand((B, 1, 28, 28), device="cuda")
    W1 = torch.randn((20, 1, 5, 5), device="cuda") * 0.28; B1 = torch.zeros((20, 1, 1), device="cuda")
    W2 = torch.randn((50, 20, 5, 5), device="cuda") * 0.063; B2 = torch.zeros((50, 1, 1), device="cuda")
    W3 = torch.randn((500, 800), device="cuda") * 0.05; B3 = torch.zeros((1, 500), device="cuda")
    W4 = torch.randn((10, 500), device="cuda") * 0.063; B4 = torch.zeros((1, 10), device="cuda")

    def step():
        C1 = am.conv2d(X, W1, B1); R1 = am.relu(C1); I1, P1 = am.maxpool2d(R1, (2, 2), (0, 0), (2, 2))
        C2 = am.conv2d(P1, W2, B2); R2 = am.relu(C2); I2, P2 = am.maxpool2d(R2, (2, 2), (0, 0), (2, 2))
        F = P2.reshape(B, 800)
        H = am.linear(F, W3, B3); RH = am.relu(H); LG = am.linear(RH, W4, B4)
        loss = am.sparse_softmax_cross_entropy_dev(LG, lab)
        GL = am.sparse_softmax_cross_entropy_backward(1.0, LG, lab)
        GRH, GW4, GB4 = am.linear_backward(RH, W4, GL)
        GH = am.relu_backward(GRH, H)
        GF, GW3, GB3 = am.linear_backward(F, W3, GH)
        GC2 = am.relu_backward(am.maxpool2d_backward(R2.shape, I2, GF.reshape(P2.shape), windows_overlap=False), C2)
        GP1, GW2, GB2 = am.conv2d_backward(P1, W2, B2, (0, 0), (1, 1), (1, 1), GC2)
        GC1 = am.relu_backward(am.maxpool2d_backward(R1.shape, I1, GP1, windows_overlap=False), C1)
        am.conv2d_backward(X, W1, B1, (0, 0), (1, 1), (1, 1), GC1, need_input_grad=False)
        return loss
    med, best = _time_gpu(step, reps=5)
    res["lenet_step_b4096"] = {"ms": med, "ms_best": best, "images_per_s": B / med * 1e3}
    return res


EXPERIMENTS = ["peaks", "nn_speed", "peaks_small_n", "simt_parity", "conv_parity", "tc1_f2", "tc2_f1", "tc2_f2", "tc2_f4", "tc2_f8", "simt_speed", "f64_speed",
               "tc1_speed_f2", "tc2_speed_f1", "tc2_speed_f2", "tc2_speed_f4", "tc2_speed_f8", "conv_speed"]


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) >= 3 and sys.argv[1] == "one":
        name = sys.argv[2]
        r = globals()[f"exp_{name}"]()
        print("RESULT " + json.dumps({"exp": name, "result": r}))
        return
    names = sys.argv[1:] or EXPERIMENTS
    with open(os.path.join(OUT, "bringup.jsonl"), "a") as f:
        for name in names:
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "one", name], capture_output=True,
                                   text=True, timeout=420)
                line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                rec = json.loads(line[-1][7:]) if line else {"exp": name, "error": (p.stdout[-1500:] + p.stderr[-3000:])}
                rec["rc"] = p.returncode
            except subprocess.TimeoutExpired:
                rec = {"exp": name, "error": "TIMEOUT"}
            rec["secs"] = round(time.time() - t0, 1)
            f.write(json.dumps(rec) + "\n")
            f.flush()
            print(json.dumps(rec)[:3000], flush=True)


if __name__ == "__main__":
    main()
