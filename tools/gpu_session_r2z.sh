#!/bin/bash
AM_B200_LIB=$PWD/arraymancer_b200/libarraymancer_b200_prof.so python - <<'PY' 2>&1 | grep "wgrad_tc dbg\|mode" | tee gpurun_out/r2z_waits.txt
import sys; sys.path.insert(0, '.')
import torch, arraymancer_b200 as am
from arraymancer_b200 import _capi
X = torch.rand((4096, 20, 12, 12), device="cuda"); W = torch.randn((50, 20, 5, 5), device="cuda") * 0.06; B = torch.rand((50, 1, 1), device="cuda")
G = torch.rand((4096, 50, 8, 8), device="cuda") - 0.5
for _ in range(2): am.conv2d_backward(X, W, B, (0, 0), (1, 1), (1, 1), G, need_input_grad=False)
torch.cuda.synchronize()
for dbg in (1, 7, 3, 5):
    _capi.set_tuning("convtc_debug", dbg)
    print("mode skip bits", dbg >> 1, file=sys.stderr, flush=True)
    am.conv2d_backward(X, W, B, (0, 0), (1, 1), (1, 1), G, need_input_grad=False); torch.cuda.synchronize()
PY
python - <<'PY' 2>&1 | grep -v "dbg\]" | tee gpurun_out/r2z_time.txt
import sys; sys.path.insert(0, '.')
import torch, arraymancer_b200 as am
from arraymancer_b200 import _capi
fl = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        fl.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return round(1e3 * ts[len(ts) // 2], 1)
X = torch.rand((4096, 20, 12, 12), device="cuda"); W = torch.randn((50, 20, 5, 5), device="cuda") * 0.06; B = torch.rand((50, 1, 1), device="cuda")
G = torch.rand((4096, 50, 8, 8), device="cuda") - 0.5
for dbg in (0, 2, 4, 6):
    _capi.set_tuning("convtc_debug", dbg)
    print("skip bits", dbg >> 1, "cv2 wgrad-only us", t(lambda: am.conv2d_backward(X, W, B, (0, 0), (1, 1), (1, 1), G, need_input_grad=False)))
_capi.set_tuning("convtc_debug", 0)
PY
