"""Per-role wait-cycle counters of the tcgen05 conv kernel (AM_CONVTC_DEBUG=1): where does a CTA's time go?"""
import os, sys
os.environ["AM_CONVTC_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import arraymancer_b200 as am
am._capi.set_conv_path(am._capi.CONV_TC)
for xs, ks in [((4096, 1, 28, 28), (20, 1, 5, 5)), ((4096, 20, 12, 12), (50, 20, 5, 5))]:
    X = torch.rand(xs, device="cuda"); W = torch.randn(ks, device="cuda") * 0.1; B = torch.zeros(ks[0], 1, 1, device="cuda")
    for _ in range(2):
        print("fwd", xs, file=sys.stderr); out = am.conv2d(X, W, B); torch.cuda.synchronize()
    print("dgrad", xs, file=sys.stderr)
    am.conv2d_backward(X, W, B, (0, 0), (1, 1), (1, 1), torch.ones_like(out), need_kernel_grad=False); torch.cuda.synchronize()
    print("wgrad", xs, file=sys.stderr)
    am.conv2d_backward(X, W, B, (0, 0), (1, 1), (1, 1), torch.ones_like(out), need_input_grad=False); torch.cuda.synchronize()
