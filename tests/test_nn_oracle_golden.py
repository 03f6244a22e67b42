"""CPU: the numpy restatement of the small NN primitives (oracle/nn_oracle.py) against the reference's own vectors
(tests/nn_primitives/test_nnp_maxpool.nim, test_nnp_loss.nim) and against independent formulas."""
import numpy as np

from oracle import nn_oracle as O
from tests.golden import known_answers as KA


def test_maxpool_reference_vectors():
    c = KA.MAXPOOL
    a = np.array(c["input"], dtype=np.int64).reshape(1, 1, 4, 4)
    idx, pooled = O.maxpool2d(a, c["kernel"], c["padding"], c["stride"])
    assert pooled.reshape(-1).tolist() == c["maxpooled"]
    assert idx.tolist() == c["max_indices"]
    # test_nnp_maxpool.nim:34-43: backward == input * numerical gradient of sum(maxpool): 1 at the argmax positions
    grad = O.maxpool2d_backward(a.shape, idx, pooled.astype(np.float64))
    want = np.zeros(16); want[c["max_indices"]] = c["maxpooled"]
    assert np.array_equal(grad.reshape(-1), want)


def test_maxpool_first_maximum_and_padding():
    x = np.array([[2., 2.], [2., 2.]]).reshape(1, 1, 2, 2)
    idx, pooled = O.maxpool2d(x, (2, 2), (1, 1), (1, 1))            # every window has padding; ties -> first in scan order
    assert pooled.shape == (1, 1, 3, 3) and np.all(pooled == 2)
    assert idx.tolist() == [0, 0, 1, 0, 0, 1, 2, 2, 3]
    g = O.maxpool2d_backward(x.shape, idx, np.arange(1., 10.).reshape(1, 1, 3, 3))
    assert g.reshape(-1).tolist() == [5., 6., 8., 9.]               # assignment: the last writer wins


def test_softmax_ce_reference_vector_and_gradient():
    c = KA.SOFTMAX_CE
    pred = np.array(c["predicted"], dtype=np.float64)
    lab = np.array(c["sparse_truth"])
    loss = O.sparse_softmax_cross_entropy(pred, lab)
    assert abs(loss - c["loss"]) <= c["tol"]
    # analytic backward (scaled by the loss, as the reference test does) vs central differences
    g = O.sparse_softmax_cross_entropy_backward(loss, pred, lab)
    num = np.zeros_like(pred)
    h = 1e-6
    for j in range(pred.shape[1]):
        p1 = pred.copy(); p1[0, j] += h
        p2 = pred.copy(); p2[0, j] -= h
        num[0, j] = (O.sparse_softmax_cross_entropy(p1, lab) - O.sparse_softmax_cross_entropy(p2, lab)) / (2 * h)
    want = loss * num
    mre = np.mean(np.abs(g - want) / np.maximum(np.abs(g), np.abs(want)))
    assert mre < 1e-5


def test_softmax_ce_batch_matches_dense_formula():
    rng = np.random.default_rng(1234)
    pred = rng.uniform(-1, 1, (256, 20))
    lab = rng.integers(0, 20, 256)
    loss = O.sparse_softmax_cross_entropy(pred, lab)
    lse = np.log(np.exp(pred).sum(axis=1))
    assert abs(loss - np.mean(lse - pred[np.arange(256), lab])) < 1e-12
    g = O.sparse_softmax_cross_entropy_backward(1.0, pred, lab)
    sm = np.exp(pred) / np.exp(pred).sum(axis=1, keepdims=True)
    sm[np.arange(256), lab] -= 1
    assert np.allclose(g, sm / 256, atol=1e-15)


def test_relu_and_linear():
    x = np.array([-1.0, 0.0, 2.5, np.nan, -0.0])
    y = O.relu(x)
    assert y[0] == 0 and y[1] == 0 and y[2] == 2.5 and np.isnan(y[3]) and y[4] == 0
    gb = O.relu_backward(np.ones(5), x)
    assert gb.tolist()[:3] == [0.0, 0.0, 1.0] and gb[3] == 1.0 and gb[4] == 0.0     # NaN cached: comparison false -> gradient
    rng = np.random.default_rng(0)
    X = rng.random((5, 7)); W = rng.random((3, 7)); b = rng.random((1, 3)); gO = rng.random((5, 3))
    assert np.allclose(O.linear(X, W, b), X @ W.T + b)
    gi, gw, gbias = O.linear_backward(X, W, gO)
    assert np.allclose(gi, gO @ W) and np.allclose(gw, gO.T @ X) and np.allclose(gbias, gO.sum(0))
