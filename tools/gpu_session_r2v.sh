#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/r2v_time.txt
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
for hr in (1, 0):
    _capi.set_tuning("convtc_hi_resident", hr)
    y = am.conv2d(X, W, B)
    print("hi_resident", hr, "cv2 fwd us", t(lambda: am.conv2d(X, W, B)), "rel", float((y.double() - ref).norm() / ref.norm()))
_capi.set_tuning("convtc_hi_resident", 1)
PY
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nn_ops.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2v_pytest.txt
