"""Worker of tests/test_gpu_multirank.py: launched by torchrun with one process per GPU (world >= 2).  Checks, against
single-GPU results computed on every rank and against the CPU oracle on sampled rows, the one-process-per-GPU paths of
arraymancer_b200.distributed: RowShardedGemm (fused epilogue all-gather and NCCL all-gather, float32 / float64 / int64),
HostShardedGemmF32 (host buffers, K slices of B exchanged over NVLink) and the batch-sharded conv2d forward / backward.
Prints 'MG_OK <rank>' on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arraymancer_b200 as am  # noqa: E402
from arraymancer_b200 import distributed as D  # noqa: E402


def rel(a, b):
    a = a.double(); b = b.double()
    return float((a - b).norm() / b.norm())


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from oracle import laser_oracle as orc
    orc.build()

    # ---- row-sharded GEMM, every rank builds the same full operands from a seed
    M, N, K = 2048 * world, 1536, 1024
    g = torch.Generator(device=dev); g.manual_seed(11)
    A = torch.rand((M, K), device=dev, generator=g) * 2 - 1
    B = torch.rand((K, N), device=dev, generator=g) * 2 - 1
    ref = torch.empty((M, N), device=dev)
    am.gemm_strided(1, A, B, 0, ref)
    for fused in (True, False):
        chunks = 1 if fused else 2
        sg = D.RowShardedGemm(M, N, K, torch.float32, dev, chunks=chunks, fused=fused)
        A_loc = D.shard_rows(A, world, rank, chunks)
        for _ in range(2):                                   # second call exercises the repack path + entry barrier
            C = sg(A_loc, B)
        torch.cuda.synchronize(); dist.barrier()
        assert rel(C, ref) <= 1e-6, ("f32 row-sharded", fused, rel(C, ref))
    rows = np.sort(np.random.default_rng(rank).choice(M, 8, replace=False))
    want = orc.matmul(np.ascontiguousarray(A[torch.from_numpy(rows).to(dev)].cpu().numpy()), B.cpu().numpy())
    got = C[torch.from_numpy(rows).to(dev)].cpu().numpy()
    assert np.linalg.norm(got.astype(np.float64) - want) / np.linalg.norm(want.astype(np.float64)) <= 5e-6

    for dt in (torch.float64, torch.int64):
        if dt == torch.int64:
            A2 = torch.randint(-2**62, 2**62, (512 * world, 300), device=dev, dtype=dt, generator=g)
            B2 = torch.randint(-2**62, 2**62, (300, 260), device=dev, dtype=dt, generator=g)
        else:
            A2 = torch.rand((512 * world, 300), device=dev, dtype=dt, generator=g); B2 = torch.rand((300, 260), device=dev, dtype=dt, generator=g)
        ref2 = torch.empty((A2.shape[0], 260), device=dev, dtype=dt)
        am.gemm_strided(1, A2, B2, 0, ref2)
        sg = D.RowShardedGemm(A2.shape[0], 260, 300, dt, dev, chunks=2)
        C2 = sg(D.shard_rows(A2, world, rank, 2), B2)
        torch.cuda.synchronize(); dist.barrier()
        assert torch.equal(C2, ref2), ("row-sharded NCCL must equal the single-GPU kernel result", dt)

    # ---- host-buffer sharded GEMM: rank's rows of A, its parts of B's K slices, rows of C
    Mh, Nh, Kh, S = 1024 * world, 1280, 256 * world * 4, 4
    Ah = torch.rand((Mh, Kh), device=dev, generator=g) * 2 - 1
    Bh = torch.rand((Kh, Nh), device=dev, generator=g) * 2 - 1
    refh = torch.empty((Mh, Nh), device=dev)
    am.gemm_strided(1, Ah, Bh, 0, refh)
    rl = Mh // world
    hA = Ah[rank * rl:(rank + 1) * rl].cpu().pin_memory()
    kc = Kh // S; part = kc // world
    hB = torch.cat([Bh[c * kc + rank * part:c * kc + (rank + 1) * part] for c in range(S)]).cpu().pin_memory()
    hC = torch.empty((rl, Nh)).pin_memory()
    hs = D.HostShardedGemmF32(rl, Nh, Kh, dev, slices=S)
    for _ in range(2):
        hs(hA, hB, hC)
    assert rel(hC.to(dev), refh[rank * rl:(rank + 1) * rl]) <= 2e-6, ("host-sharded", rel(hC.to(dev), refh[rank * rl:(rank + 1) * rl]))

    # ---- batch-sharded conv2d fwd / bwd
    NB = 64 * world
    X = torch.rand((NB, 20, 12, 12), device=dev, generator=g); W = torch.randn((50, 20, 5, 5), device=dev, generator=g) * 0.06
    Bv = torch.rand((50, 1, 1), device=dev, generator=g); G = torch.rand((NB, 50, 8, 8), device=dev, generator=g) - 0.5
    out_ref = am.conv2d(X, W, Bv)
    gi_ref, gw_ref, gb_ref = am.conv2d_backward(X, W, Bv, (0, 0), (1, 1), (1, 1), G)
    lo, hi = D.shard_batch(NB, world, rank)
    out = D.conv2d_batch_sharded(X[lo:hi].contiguous(), W, Bv)
    gi, gw, gb = D.conv2d_backward_batch_sharded(X[lo:hi].contiguous(), W, Bv, (0, 0), (1, 1), (1, 1), G[lo:hi].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(out, out_ref[lo:hi]) and rel(gi, gi_ref[lo:hi]) <= 1e-6
    assert rel(gw, gw_ref) <= 1e-5 and rel(gb, gb_ref) <= 1e-5, (rel(gw, gw_ref), rel(gb, gb_ref))
    dist.barrier()
    print(f"MG_OK {rank}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
