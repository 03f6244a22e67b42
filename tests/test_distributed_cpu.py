"""world_size-2 gloo tests of the multi-GPU host logic (block-cyclic row sharding + all-gather of C,
batch-sharded conv with all-reduced weight gradients).  The per-rank compute is the CPU oracle here
(injected); on GPUs the same code runs the CUDA kernels over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from arraymancer_b200 import distributed as D
        from oracle import laser_oracle as orc

        def local_gemm(A, B, C):
            C.copy_(torch.from_numpy(orc.matmul(A.numpy(), B.numpy())))

        rng = np.random.default_rng(123)
        M, N, K, chunks = 48, 20, 300, 3
        a = rng.integers(-2**63, 2**63 - 1, size=(M, K), dtype=np.int64)
        b = rng.integers(-2**63, 2**63 - 1, size=(K, N), dtype=np.int64)
        A_full, B = torch.from_numpy(a), torch.from_numpy(b)
        op = D.RowShardedGemm(M, N, K, torch.int64, "cpu", chunks=chunks, local_gemm=local_gemm)
        C = op(D.shard_rows(A_full, world, rank, chunks), B)
        want = orc.matmul(a, b)
        ok_gemm = bool(np.array_equal(C.numpy(), want))      # bit-exact under sharding (no K split)

        # ownership covers every row exactly once
        rows = sorted(r for rk in range(world) for lo, hi in D.owned_row_ranges(M, world, rk, chunks) for r in range(lo, hi))
        ok_rows = rows == list(range(M))

        # batch-sharded conv: forward stays sharded; backward all-reduces grad_kernel / grad_bias
        x = rng.random((5, 3, 7, 6)); k = rng.random((4, 3, 3, 3)) - 0.5; bias = rng.random((4, 1, 1))
        go = rng.random((5, 4, 7, 6))
        lo, hi = D.shard_batch(5, world, rank)

        def conv_fwd(xl, kk, bb, pad, st, dil):
            return torch.from_numpy(orc.conv2d(xl.numpy(), kk.numpy(), bb.numpy(), pad, st, dil))

        def conv_bwd(xl, kk, bb, pad, st, dil, gol):
            gi, gw, gb = orc.conv2d_backward(xl.numpy(), kk.numpy(), gol.numpy(), True, pad, st, dil)
            return torch.from_numpy(gi), torch.from_numpy(gw), torch.from_numpy(gb)

        X, Kt, Bt, GO = (torch.from_numpy(v) for v in (x, k, bias, go))
        out = D.conv2d_batch_sharded(X[lo:hi], Kt, Bt, (1, 1), (1, 1), (1, 1), local_conv=conv_fwd)
        full = orc.conv2d(x, k, bias, (1, 1))
        ok_fwd = bool(np.allclose(out.numpy(), full[lo:hi], rtol=1e-13, atol=1e-13))
        gi, gw, gb = D.conv2d_backward_batch_sharded(X[lo:hi], Kt, Bt, (1, 1), (1, 1), (1, 1), GO[lo:hi],
                                                     local_conv_backward=conv_bwd)
        wgi, wgw, wgb = orc.conv2d_backward(x, k, go, True, (1, 1))
        ok_bwd = bool(np.allclose(gi.numpy(), wgi[lo:hi], rtol=1e-12, atol=1e-12) and
                      np.allclose(gw.numpy(), wgw, rtol=1e-12, atol=1e-12) and
                      np.allclose(gb.numpy(), wgb, rtol=1e-12, atol=1e-12))
        q.put((rank, ok_gemm, ok_rows, ok_fwd, ok_bwd))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_row_sharded_gemm_and_batch_sharded_conv_world2():
    from oracle import laser_oracle
    laser_oracle.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(r[0] for r in res) == [0, 1]
    for r in res:
        assert all(r[1:]), f"rank {r[0]}: gemm={r[1]} rows={r[2]} conv_fwd={r[3]} conv_bwd={r[4]}"


def test_shard_helpers():
    from arraymancer_b200 import distributed as D
    assert D.chunk_rows(32768, 8, 4) == 1024
    with pytest.raises(ValueError):
        D.chunk_rows(100, 8, 4)
    spans = [D.shard_batch(4096, 8, r) for r in range(8)]
    assert spans[0] == (0, 512) and spans[-1] == (3584, 4096)
    spans = [D.shard_batch(10, 4, r) for r in range(4)]
    assert spans == [(0, 3), (3, 6), (6, 8), (8, 10)]
