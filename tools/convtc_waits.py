"""Per-role wait-cycle breakdown of the tcgen05 conv kernels (cv2 shape, batch 4096) from the profiling build:
   make -C arraymancer_b200/csrc prof && AM_B200_LIB=arraymancer_b200/libarraymancer_b200_prof.so python tools/convtc_waits.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, arraymancer_b200 as am
from arraymancer_b200 import _capi
X = torch.rand((4096, 20, 12, 12), device="cuda"); W = torch.randn((50, 20, 5, 5), device="cuda") * 0.06
B = torch.rand((50, 1, 1), device="cuda"); G = torch.rand((4096, 50, 8, 8), device="cuda") - 0.5
for _ in range(2):
    am.conv2d(X, W, B); am.conv2d_backward(X, W, B, (0, 0), (1, 1), (1, 1), G)
torch.cuda.synchronize()
_capi.set_tuning("convtc_debug", 1)
for _ in range(2):
    am.conv2d(X, W, B); am.conv2d_backward(X, W, B, (0, 0), (1, 1), (1, 1), G)
torch.cuda.synchronize()
