"""Times the LeNet conv layers (batch 4096, L2 flushed before every repetition): usage time_conv.py [cv1|cv2|all]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, arraymancer_b200 as am
what = sys.argv[1] if len(sys.argv) > 1 else "all"
fl = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, reps=30):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        fl.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return round(1e3 * ts[len(ts) // 2], 1)
layers = {"cv1": ((4096, 1, 28, 28), (20, 1, 5, 5)), "cv2": ((4096, 20, 12, 12), (50, 20, 5, 5))}
for name, (xs, ks) in layers.items():
    if what not in ("all", name): continue
    X = torch.rand(xs, device="cuda"); W = torch.randn(ks, device="cuda") * 0.1; B = torch.rand((ks[0], 1, 1), device="cuda")
    y = am.conv2d(X, W, B); G = torch.rand_like(y) - 0.5
    ref = torch.nn.functional.conv2d(X.double(), W.double(), B.double().reshape(-1))
    print(name, "fwd us", t(lambda: am.conv2d(X, W, B)), "rel", float((y.double() - ref).norm() / ref.norm()))
    gi, gw, gb = am.conv2d_backward(X, W, B, (0, 0), (1, 1), (1, 1), G)
    Xd = X.double().requires_grad_(True); Wd = W.double().requires_grad_(True)
    torch.nn.functional.conv2d(Xd, Wd).backward(G.double())
    print(name, "bwd us", t(lambda: am.conv2d_backward(X, W, B, (0, 0), (1, 1), (1, 1), G)), "rel gi", float((gi.double() - Xd.grad).norm() / Xd.grad.norm()),
          "gw", float((gw.double() - Wd.grad).norm() / Wd.grad.norm()), "gb", float((gb.double().flatten() - G.double().sum((0, 2, 3))).norm() / G.double().sum((0, 2, 3)).norm()))
