"""CPU restatement (numpy) of the reference's small NN primitives around the contractions — TEST INFRASTRUCTURE ONLY
(imported by tests/ and __graft_entry__.smoke(); never by the product path).

Each function follows the cited reference lines (paths relative to /root/reference/src/arraymancer/nn_primitives/):
  relu, relu_backward                      nnp_activation.nim:35-36, 65-70
  maxpool2d, maxpool2d_backward            nnp_maxpooling.nim:19-83
  linear, linear_backward                  nnp_linear.nim:20-66
  sparse_softmax_cross_entropy (+backward) nnp_softmax_cross_entropy.nim:100-178, 219-252; private/p_logsumexp.nim:13-23
Pinned by the reference's own vectors: tests/nn_primitives/test_nnp_maxpool.nim:21-32 and test_nnp_loss.nim:29-44
(tests/golden/known_answers.py MAXPOOL / SOFTMAX_CE)."""
import numpy as np

INT_LOW = np.iinfo(np.int64).min


def relu(x):
    # t.map_inline max(0.T, x); Nim's max(a, b) returns b unless b <= a: NaN stays NaN
    return np.where(x <= 0, np.zeros((), x.dtype), x)


def relu_backward(gradient, cached):
    # if x <= 0.T: 0.T else: y
    return np.where(cached <= 0, np.zeros((), gradient.dtype), gradient)


def maxpool2d(x, kernel, padding=(0, 0), stride=(1, 1)):
    """-> (max_indices [N*C*outH*outW] int64 flat input indices, maxpooled [N,C,outH,outW])"""
    N, C, H, W = x.shape
    kH, kW = kernel
    outH = (H + 2 * padding[0] - kH) // stride[0] + 1
    outW = (W + 2 * padding[1] - kW) // stride[1] + 1
    low = -np.inf if np.issubdtype(x.dtype, np.floating) else np.iinfo(x.dtype).min
    best = np.full((N, C, outH, outW), low, dtype=x.dtype)
    arg = np.full((N, C, outH, outW), INT_LOW, dtype=np.int64)
    flat = np.arange(N * C * H * W, dtype=np.int64).reshape(N, C, H, W)
    hs = np.arange(outH) * stride[0] - padding[0]
    ws = np.arange(outW) * stride[1] - padding[1]
    for ph in range(kH):                      # window scanned in (ph, pw) order, strict '>' keeps the first maximum
        rows = hs + ph
        rok = (rows >= 0) & (rows < H)
        for pw in range(kW):
            cols = ws + pw
            cok = (cols >= 0) & (cols < W)
            rr = np.clip(rows, 0, H - 1); cc = np.clip(cols, 0, W - 1)
            v = x[:, :, rr][:, :, :, cc]
            fi = flat[:, :, rr][:, :, :, cc]
            ok = (rok[:, None] & cok[None, :])[None, None]
            with np.errstate(invalid="ignore"):
                take = ok & (v > best)
            best = np.where(take, v, best)
            arg = np.where(take, fi, arg)
    return arg.reshape(-1), best


def maxpool2d_backward(input_shape, max_indices, grad_output):
    # result = zeros; for i ascending: result[max_indices[i]] = gradOutput[i]   (assignment: the last i wins)
    out = np.zeros(int(np.prod(input_shape)), dtype=grad_output.dtype)
    go = grad_output.reshape(-1)
    ok = max_indices >= 0
    out[max_indices[ok]] = go[ok]             # numpy fancy assignment keeps the last occurrence
    return out.reshape(input_shape)


def linear(x, weight, bias=None):
    y = x @ weight.T
    if bias is not None:
        y = y + bias.reshape(1, -1)
    return y


def linear_backward(x, weight, grad_output, with_bias=True):
    gi = grad_output @ weight
    gw = grad_output.T @ x
    gb = grad_output.sum(axis=0) if with_bias else None
    return gi, gw, gb


def _stream_max_sumexp(row):
    m = -np.inf
    s = row.dtype.type(0)
    one = row.dtype.type(1)
    for v in row:
        if v <= m:
            s = s + np.exp(v - m)
        else:
            s = s * np.exp(m - v) + one if np.isfinite(m) else one
            m = v
    return m, s


def sparse_softmax_cross_entropy(x, labels):
    batch = x.shape[0]
    if batch == 0:
        return x.dtype.type(0)
    total = x.dtype.type(0)
    for i in range(batch):
        m, s = _stream_max_sumexp(x[i])
        total += np.log(s) + m - x[i, int(labels[i])]
    return total / x.dtype.type(batch)


def sparse_softmax_cross_entropy_backward(gradient, cached, labels):
    batch = cached.shape[0]
    out = np.zeros_like(cached)
    out[np.arange(batch), labels.astype(np.int64)] = -1
    for i in range(batch):
        m, s = _stream_max_sumexp(cached[i])
        out[i] = gradient * (np.exp(cached[i] - m) / s + out[i]) / cached.dtype.type(batch)
    return out
