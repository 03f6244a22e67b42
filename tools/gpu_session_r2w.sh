#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/r2w_time.txt
import sys; sys.path.insert(0, '.')
import torch, arraymancer_b200 as am
from arraymancer_b200 import _capi
fl = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, reps=30):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        fl.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return round(1e3 * ts[len(ts) // 2], 1)
X = torch.rand((4096, 20, 12, 12), device="cuda"); W = torch.randn((50, 20, 5, 5), device="cuda") * 0.06; B = torch.rand((50, 1, 1), device="cuda")
ref = torch.nn.functional.conv2d(X.double(), W.double(), B.double().reshape(-1))
G = torch.rand((4096, 50, 8, 8), device="cuda") - 0.5
Xd = X.double().requires_grad_(True); Wd = W.double().requires_grad_(True)
torch.nn.functional.conv2d(Xd, Wd).backward(G.double())
for fk in (2, 4, 8, 16):
    _capi.set_tuning("tc_flush_kb", fk)
    y = am.conv2d(X, W, B)
    gi, gw, gb = am.conv2d_backward(X, W, B, (0, 0), (1, 1), (1, 1), G)
    print("flush", fk, "cv2 fwd us", t(lambda: am.conv2d(X, W, B)), "rel", float((y.double() - ref).norm() / ref.norm()),
          "| bwd us", t(lambda: am.conv2d_backward(X, W, B, (0, 0), (1, 1), (1, 1), G)), "rel gi", float((gi.double() - Xd.grad).norm() / Xd.grad.norm()), "gw", float((gw.double() - Wd.grad).norm() / Wd.grad.norm()))
_capi.set_tuning("tc_flush_kb", 2)
PY
