"""Multi-GPU sharding of the hot path — one process per GPU, torch.distributed for the plumbing.

The reference has no multi-device code at all; parity is defined against the
single-device result:

  * Large GEMM: rows of A (and of C) are dealt to the ranks in block-cyclic chunks, B is replicated.
    Rank r owns, for every chunk index j, the rows [(j*g + r)*mc, (j*g + r + 1)*mc).  With that
    layout the all-gather of chunk j over the ranks fills the CONTIGUOUS row range
    [j*g*mc, (j+1)*g*mc) of the full row-major C, so every chunk is one in-place
    `all_gather_into_tensor` — issued on a communication stream while the GEMM of chunk j+1 runs
    (compute of chunk j+1 overlaps the NVLink transfer of chunk j).  No K split, so integer results
    stay bit-exact and float results are identical to the 1-GPU result.
  * Batched conv2d: images are independent — the batch is split over the ranks, weights are
    replicated; backward all-reduces grad_kernel / grad_bias (sum), grad_input stays sharded.

`local_gemm` / `local_conv*` are injectable so the host-side logic (partitioning + collectives) is
testable with the gloo backend on CPU (tests/test_distributed_cpu.py plugs the oracle in there);
on GPUs they default to the CUDA kernels.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def chunk_rows(M: int, world: int, chunks: int) -> int:
    """Rows per (rank, chunk) block; M must split evenly (pad the operand otherwise)."""
    if M % (world * chunks) != 0:
        raise ValueError(f"row-sharded GEMM: M={M} must be a multiple of world*chunks={world * chunks}")
    return M // (world * chunks)


def owned_row_ranges(M: int, world: int, rank: int, chunks: int):
    """Global row ranges [(start, stop)] owned by `rank`, in local storage order."""
    mc = chunk_rows(M, world, chunks)
    return [((j * world + rank) * mc, (j * world + rank + 1) * mc) for j in range(chunks)]


def shard_rows(A_full: torch.Tensor, world: int, rank: int, chunks: int) -> torch.Tensor:
    """Pick this rank's block-cyclic rows out of a full A (test / setup helper)."""
    return torch.cat([A_full[a:b] for a, b in owned_row_ranges(A_full.shape[0], world, rank, chunks)], dim=0)


class RowShardedGemm:
    """C[M,N] = A[M,K] @ B[K,N] with A's rows dealt over the ranks and the full C gathered on every
    rank.  `A_local` holds this rank's rows ([chunks*mc, K], chunk-major); `B` is replicated."""

    def __init__(self, M: int, N: int, K: int, dtype, device, chunks: int = 4, group=None, local_gemm=None,
                 fused: bool = False):
        """fused=True (float32, CUDA, world > 1): C lives in symmetric memory and the GEMM epilogue stores every
        tile to all GPUs' copies (am_gemm_packed_f32_bcast) — no NCCL collective, one launch per chunk."""
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.M, self.N, self.K, self.chunks = M, N, K, chunks
        self.mc = chunk_rows(M, self.world, chunks)
        self.cuda = torch.device(device).type == "cuda"
        self.sym = None
        if fused and self.cuda and self.world > 1:
            if dtype != torch.float32:
                raise TypeError("RowShardedGemm(fused=True) is the float32 tcgen05 path")
            self.sym = SymmetricResult((M, N), dtype, device, group)
            self.C = self.sym.C
            self._pA = self._pB = None
        else:
            self.C = torch.empty((M, N), dtype=dtype, device=device)  # full result, row-major
        if local_gemm is None:
            from .cuda_tensor import gemm_strided
            local_gemm = lambda A, B, C: gemm_strided(1, A, B, 0, C)  # noqa: E731
        self.local_gemm = local_gemm
        self.comm_stream = torch.cuda.Stream(device=device) if self.cuda else None

    def _call_fused(self, A_local: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
        from .cuda_tensor import PackedF32, gemm_packed_bcast
        g, r, mc = self.world, self.rank, self.mc
        # cross-GPU ordering: a fast rank must not store call n+1's tiles into a peer's copy of C while that peer
        # (or a consumer it launched) still reads call n's C — every rank passes this barrier only after its earlier
        # work on the current stream, reads of C included, has been enqueued ahead of it
        self.sym.barrier()
        first = self._pB is None
        if first:
            self._pB = PackedF32(B, "b")
            self._pA = PackedF32(A_local[:mc], "a")          # packs chunk 0
        else:
            self._pB.repack(B)
        for j in range(self.chunks):
            lo = (j * g + r) * mc
            mine = self.C[lo:lo + mc]
            if not (first and j == 0):
                self._pA.repack(A_local[j * mc:(j + 1) * mc])
            gemm_packed_bcast(1.0, self._pA, self._pB, mine, self.sym.peer_ptrs(mine), self.sym.rank)
        self.sym.barrier()
        return self.C

    def __call__(self, A_local: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
        if self.sym is not None:
            return self._call_fused(A_local, B)
        g, r, mc = self.world, self.rank, self.mc
        works = []
        for j in range(self.chunks):
            lo = (j * g + r) * mc
            mine = self.C[lo:lo + mc]
            self.local_gemm(A_local[j * mc:(j + 1) * mc], B, mine)
            if g == 1:
                continue
            span = self.C[j * g * mc:(j + 1) * g * mc]
            if self.cuda:
                ready = torch.cuda.Event()
                ready.record()
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(ready)
                    works.append(dist.all_gather_into_tensor(span, mine, group=self.group, async_op=True))
            else:
                parts = [span[i * mc:(i + 1) * mc] for i in range(g)]
                dist.all_gather(parts, mine.clone(), group=self.group)
        for w in works:
            w.wait()
        if self.cuda and g > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        return self.C


class SymmetricResult:
    """The full result matrix C allocated as symmetric memory (torch.distributed._symmetric_memory:
    one identically-sized allocation per GPU, each mapped into every process of the node over
    NVLink / NVSwitch).  `peer_ptrs(view)` gives, for a view of this rank's C, the address of the
    same block in every GPU's copy — what am_gemm_packed_f32_bcast stores to, so the all-gather of
    SURVEY 8e happens inside the GEMM epilogue instead of as a separate NCCL collective."""

    def __init__(self, shape, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.C = symm.empty(*shape, dtype=dtype, device=device)
        self.handle = symm.rendezvous(self.C, group)
        self.world = self.handle.world_size
        self.rank = self.handle.rank
        self.bases = [int(x) for x in self.handle.buffer_ptrs]
        if len(self.bases) != self.world or any(b == 0 for b in self.bases):
            raise RuntimeError("symmetric memory: peer mapping unavailable")
        self._base = self.C.data_ptr()

    def peer_ptrs(self, view: torch.Tensor):
        off = view.data_ptr() - self._base
        return [b + off for b in self.bases]

    def barrier(self):
        """Device-side barrier on the current stream across the ranks: afterwards every rank's
        earlier stores to the peers are complete and visible."""
        self.handle.barrier()


class HostShardedGemmF32:
    """C_local[rows, N] = A_local[rows, K] @ B[K, N] from HOST (pinned) buffers on every rank of a row-sharded float32
    GEMM — the end-to-end twin of RowShardedGemm.  Every rank holds its rows of A, and of B only its share of every K
    slice: B is cut into `slices` row blocks of K/slices rows, each dealt over the ranks in g contiguous parts, and
    `B_part` ([K/g, N], pinned) holds this rank's parts, slice after slice.  Per K slice c:
        upload stream   A_local[:, slice c] (pitched copy) and this rank's part of B's slice c, into symmetric memory
        exchange        cross-GPU barrier, then the g-1 other parts are PULLED from the peers' symmetric buffers by
                        the copy engines over NVLink (no SMs, so the persistent GEMM keeps all of them)
        compute stream  split/pack of both slices, C_local (+)= A_c @ B_c on the tcgen05 path
    and the last slice is multiplied by row chunks so that finished rows of C go back to the host while the rest
    computes.  Total H2D over all ranks = |A| + |B| (it was |A| + g*|B| when every rank uploaded all of B)."""

    def __init__(self, rows_local: int, N: int, K: int, device, group=None, slices: int = 4, row_chunks: int = 4):
        import torch.distributed._symmetric_memory as symm
        from .cuda_tensor import PackedF32
        self.group = group if group is not None else dist.group.WORLD
        self.g, self.r = dist.get_world_size(group), dist.get_rank(group)
        g = self.g
        if K % (slices * g * 32) != 0:
            raise ValueError(f"HostShardedGemmF32: K={K} must be a multiple of slices*world*32 = {slices * g * 32}")
        self.rows, self.N, self.K, self.S = rows_local, N, K, slices
        self.kc = K // slices                  # rows of B per K slice
        self.part = self.kc // g               # ... of which this rank uploads `part`
        self.dev = torch.device(device)
        self.dA = torch.empty((rows_local, K), dtype=torch.float32, device=device)
        self.dC = torch.empty((rows_local, N), dtype=torch.float32, device=device)
        # two K slices of B in flight (double buffer), in symmetric memory so peers can read the parts
        self.symB = symm.empty(2 * self.kc * N, dtype=torch.float32, device=device)
        self.hB = symm.rendezvous(self.symB, self.group)
        self.peerB = [int(x) for x in self.hB.buffer_ptrs]
        if len(self.peerB) != g or any(b == 0 for b in self.peerB):
            raise RuntimeError("symmetric memory: peer mapping unavailable")
        self.s_in = torch.cuda.Stream(device=device)
        self.s_x = torch.cuda.Stream(device=device)      # exchange (barrier + peer pulls)
        self.s_out = torch.cuda.Stream(device=device)
        self.pA = PackedF32(self.dA[:, :self.kc], "a")
        self.pB = PackedF32(self.symB[:self.kc * N].view(self.kc, N), "b")
        self.rc = max(1, row_chunks)
        self.pAr = None
        if self.rc > 1 and rows_local % self.rc == 0 and rows_local // self.rc >= 256:
            self.pAr = PackedF32(self.dA[:rows_local // self.rc, :self.kc], "a")
        else:
            self.rc = 1

    def _copy2d(self, stream, dst_ptr, dpitch, src_ptr, spitch, width, rows, kind):
        from . import _capi
        _capi.check(_capi.lib().am_memcpy2d_async(stream.cuda_stream, dst_ptr, dpitch, src_ptr, spitch, width, rows, kind))

    def __call__(self, hA: torch.Tensor, hB_part: torch.Tensor, hC: torch.Tensor) -> torch.Tensor:
        from .cuda_tensor import gemm_packed
        g, r, N, K, kc, part, rows = self.g, self.r, self.N, self.K, self.kc, self.part, self.rows
        if tuple(hA.shape) != (rows, K) or tuple(hB_part.shape) != (K // g, N) or tuple(hC.shape) != (rows, N):
            raise IndexError("HostShardedGemmF32: buffer shapes do not match the plan")
        cur = torch.cuda.current_stream(self.dev)
        for st in (self.s_in, self.s_x, self.s_out):
            st.wait_stream(cur)
        ev_up, ev_x, ev_free = [], [], [None, None]
        base = self.symB.data_ptr()
        for c in range(self.S):
            buf = c & 1
            off = buf * kc * N                                       # element offset of this slice's buffer
            with torch.cuda.stream(self.s_in):
                if ev_free[buf] is not None:
                    self.s_in.wait_event(ev_free[buf])                # the product that read this buffer has been issued and finished
                # this rank's part of slice c: rows [r*part, (r+1)*part) of the slice buffer
                dst = self.symB[off + r * part * N: off + (r + 1) * part * N]
                dst.copy_(hB_part[c * part:(c + 1) * part].reshape(-1), non_blocking=True)
                self._copy2d(self.s_in, self.dA.data_ptr() + 4 * c * kc, 4 * K, hA.data_ptr() + 4 * c * kc, 4 * K, 4 * kc, rows, 1)
                e = torch.cuda.Event(); e.record(self.s_in); ev_up.append(e)
            with torch.cuda.stream(self.s_x):
                self.s_x.wait_event(ev_up[c])
                self.hB.barrier()                                     # every rank's part of slice c is in its symmetric buffer
                for q in range(1, g):
                    src_rank = (r + q) % g
                    o = 4 * (off + src_rank * part * N)
                    self._copy2d(self.s_x, base + o, 4 * N, self.peerB[src_rank] + o, 4 * N, 4 * N, part, 3)
                self.hB.barrier()                                     # peers have finished pulling from this rank's buffer
                e = torch.cuda.Event(); e.record(self.s_x); ev_x.append(e)
            cur.wait_event(ev_x[c])
            Bc = self.symB[off:off + kc * N].view(kc, N)
            self.pB.repack(Bc)
            last = c == self.S - 1
            if not last or self.rc == 1:
                self.pA.repack(self.dA[:, c * kc:(c + 1) * kc])
                gemm_packed(1.0, self.pA, self.pB, 0.0 if c == 0 else 1.0, self.dC)
                if last:
                    e = torch.cuda.Event(); e.record(cur)
                    with torch.cuda.stream(self.s_out):
                        self.s_out.wait_event(e)
                        hC.copy_(self.dC, non_blocking=True)
            else:
                rr = rows // self.rc
                for i in range(self.rc):
                    self.pAr.repack(self.dA[i * rr:(i + 1) * rr, c * kc:(c + 1) * kc])
                    gemm_packed(1.0, self.pAr, self.pB, 0.0 if c == 0 else 1.0, self.dC[i * rr:(i + 1) * rr])
                    e = torch.cuda.Event(); e.record(cur)
                    with torch.cuda.stream(self.s_out):
                        self.s_out.wait_event(e)
                        hC[i * rr:(i + 1) * rr].copy_(self.dC[i * rr:(i + 1) * rr], non_blocking=True)
            e = torch.cuda.Event(); e.record(cur); ev_free[buf] = e
        cur.wait_stream(self.s_out)
        cur.synchronize()                                             # the result is in hC when the call returns
        return hC


def shard_batch(n_images: int, world: int, rank: int):
    """Contiguous image range of `rank` (the remainder goes to the first ranks)."""
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def conv2d_batch_sharded(x_local, kernel, bias, padding=(0, 0), strides=(1, 1), dilation=(1, 1), local_conv=None):
    """Forward on this rank's images; the output stays batch-sharded (no collective)."""
    if local_conv is None:
        from .nn_primitives import conv2d as local_conv
    return local_conv(x_local, kernel, bias, padding, strides, dilation)


def conv2d_backward_batch_sharded(x_local, kernel, bias, padding, strides, dilation, grad_out_local, group=None,
                                  local_conv_backward=None):
    """Backward on this rank's images: grad_input stays sharded, grad_kernel / grad_bias are summed
    over the ranks (all-reduce; float summation order differs from the serial reference loop, covered
    by the stated backward tolerance; integers stay exact)."""
    if local_conv_backward is None:
        from .nn_primitives import conv2d_backward as local_conv_backward
    gin, gk, gb = local_conv_backward(x_local, kernel, bias, padding, strides, dilation, grad_out_local)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(gk, op=dist.ReduceOp.SUM, group=group)
        if gb is not None:
            dist.all_reduce(gb, op=dist.ReduceOp.SUM, group=group)
    return gin, gk, gb
