"""Launch each kernel family once (after a warm-up) so `ncu` can capture it.
usage: python tools/profile_kernels.py [simt|tc|conv|skinny|all]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import arraymancer_b200 as am  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2


def gemm(dt, n, path=None):
    if path is not None:
        am.set_f32_path(path)
    if dt.is_floating_point:
        A = torch.rand(n, n, device="cuda", dtype=dt) * 2 - 1; B = torch.rand(n, n, device="cuda", dtype=dt) * 2 - 1
    else:
        A = torch.randint(0, 100, (n, n), device="cuda", dtype=dt); B = torch.randint(0, 100, (n, n), device="cuda", dtype=dt)
    C = torch.empty(n, n, device="cuda", dtype=dt)
    for _ in range(reps):
        am.gemm_strided(1, A, B, 0, C)
    torch.cuda.synchronize()


if what in ("simt", "all"):
    gemm(torch.float64, 4096)
    gemm(torch.int64, 4096)
    gemm(torch.int32, 4096)
    gemm(torch.float32, 4096, am.F32_SIMT)
    gemm(torch.int64, 1500)
if what == "i64wide":
    n = 4096
    A = torch.randint(-2**62, 2**62, (n, n), device="cuda", dtype=torch.int64); B = torch.randint(-2**62, 2**62, (n, n), device="cuda", dtype=torch.int64)
    C = torch.empty(n, n, device="cuda", dtype=torch.int64)
    for _ in range(reps):
        am.gemm_strided(1, A, B, 0, C)
    torch.cuda.synchronize()
if what in ("tc", "all"):
    gemm(torch.float32, 8192, am.F32_TC)
if what == "convtc":
    am._capi.set_conv_path(am._capi.CONV_TC)
if what in ("conv", "convtc", "all"):
    for xs, ks in [((4096, 1, 28, 28), (20, 1, 5, 5)), ((4096, 20, 12, 12), (50, 20, 5, 5))]:
        X = torch.rand(xs, device="cuda"); W = torch.randn(ks, device="cuda") * 0.1; B = torch.zeros(ks[0], 1, 1, device="cuda")
        for _ in range(reps):
            out = am.conv2d(X, W, B)
            am.conv2d_backward(X, W, B, (0, 0), (1, 1), (1, 1), torch.ones_like(out))
        torch.cuda.synchronize()
if what == "skinny":
    n = 16384
    for shape in ((1, n, n), (n, 1, n), (16, n, n), (n, 16, n)):
        M, N, K = shape
        A = torch.rand(M, K, device="cuda") * 2 - 1; B = torch.rand(K, N, device="cuda") * 2 - 1
        C = torch.empty(M, N, device="cuda")
        for _ in range(reps):
            am.gemm_strided(1, A, B, 0, C)
        torch.cuda.synchronize()
