"""CPU: host-side argument checking of the NN-primitive mirror (arraymancer_b200/nn_primitives.py) — the errors are
raised before the C ABI is entered, so they are testable without a GPU; and there is no CPU fallback to fall into."""
import pytest

torch = pytest.importorskip("torch")

import arraymancer_b200 as am  # noqa: E402


def test_cpu_tensors_are_refused():
    x = torch.zeros((2, 3))
    with pytest.raises(ValueError):
        am.relu(x)
    with pytest.raises(ValueError):
        am.relu_backward(x, x)
    with pytest.raises(ValueError):
        am.maxpool2d(torch.zeros((1, 1, 4, 4)), (2, 2))
    with pytest.raises(ValueError):
        am.linear(x, torch.zeros((4, 3)))
    with pytest.raises(ValueError):
        am.sparse_softmax_cross_entropy(x, torch.zeros((2,), dtype=torch.int64))
    with pytest.raises(ValueError):
        am.sparse_softmax_cross_entropy_backward(1.0, x, torch.zeros((2,), dtype=torch.int64))
    with pytest.raises(ValueError):
        am.conv2d(torch.zeros((1, 1, 4, 4)), torch.zeros((1, 1, 3, 3)))


def test_exported_names_mirror_the_reference():
    # nn_primitives/nnp_{activation,maxpooling,linear,softmax_cross_entropy,convolution}.nim
    for name in ("relu", "relu_backward", "maxpool2d", "maxpool2d_backward", "linear", "linear_backward",
                 "sparse_softmax_cross_entropy", "sparse_softmax_cross_entropy_backward", "conv2d", "conv2d_backward"):
        assert callable(getattr(am, name))


def test_conv_out_dims_and_pool_dims_follow_the_cpu_formulas():
    # fallback/conv.nim:90-91 and nnp_maxpooling.nim:37-38
    assert am.conv_out_dims((4096, 1, 28, 28), (20, 1, 5, 5)) == (4096, 20, 24, 24)
    assert am.conv_out_dims((1, 3, 5, 5), (2, 3, 3, 3), (1, 1), (2, 2)) == (1, 2, 3, 3)
    assert am.conv_out_dims((1, 3, 11, 10), (70, 3, 3, 3), (2, 2), (1, 1), (2, 2)) == (1, 70, 11, 10)
