"""NVLink byte counters of the fused GEMM + all-gather epilogue: ONE host process drives two GPUs through the am_mg_* entries
(no torch.distributed), so the kernels can be profiled by ncu like any single-process program:
  gpurun --gpus 2 -- ncu --metrics nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum \\
         --clock-control none -k regex:gemm_tf32x3 --csv --log-file gpurun_out/nvlink.csv python tools/nvlink_counters.py
Each GPU computes half of the rows of C (n x n float32) and stores them to its own copy AND to the peer's: expected NVLink
payload per GPU and launch = n*n*4/2 bytes transmitted and as many received."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import arraymancer_b200 as am  # noqa: F401
from arraymancer_b200.multi_gpu import MgContext
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
devs = [0, 1]
ctx = MgContext(devs)
A_local, Bs, Cs = [], [], []
for g, d in enumerate(devs):
    r0, rows = ctx.rows(n, g)
    gen = torch.Generator(device=f"cuda:{d}"); gen.manual_seed(5 + g)
    A_local.append(torch.rand((rows, n), device=f"cuda:{d}", generator=gen) * 2 - 1)
    gb = torch.Generator(device=f"cuda:{d}"); gb.manual_seed(99)
    Bs.append(torch.rand((n, n), device=f"cuda:{d}", generator=gb) * 2 - 1)
    Cs.append(torch.zeros((n, n), device=f"cuda:{d}"))
for _ in range(2):
    ctx.gemm_rowsharded(1.0, A_local, Bs, Cs)
    ctx.synchronize()
# both copies hold the same full C
same = torch.equal(Cs[0].cpu(), Cs[1].cpu())
ref = (A_local[0][:64].double() @ Bs[0].double())
rel = float((Cs[1][:64].double().cpu() - ref.cpu()).norm() / ref.cpu().norm())
print(f"n={n} copies identical: {same}; rows 0..63 of GPU 1's copy vs fp64: rel {rel:.2e}; expected NVLink payload per GPU and launch: {n * n * 2} bytes")
ctx.close()
