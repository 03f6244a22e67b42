"""Launch the LeNet conv layers (batch 4096) a few times so ncu can capture the kernels.  usage: prof_conv.py [cv1|cv2|all] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import arraymancer_b200 as am
what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
layers = {"cv1": ((4096, 1, 28, 28), (20, 1, 5, 5)), "cv2": ((4096, 20, 12, 12), (50, 20, 5, 5))}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, (xs, ks) in layers.items():
    if what not in ("all", name):
        continue
    X = torch.rand(xs, device="cuda"); W = torch.randn(ks, device="cuda") * 0.1; B = torch.rand(ks[0], 1, 1, device="cuda")
    out = am.conv2d(X, W, B); G = torch.rand_like(out)
    for _ in range(reps):
        flush.zero_()
        am.conv2d(X, W, B)
        flush.zero_()
        am.conv2d_backward(X, W, B, (0, 0), (1, 1), (1, 1), G)
    torch.cuda.synchronize()
