"""bench.py --workload dgemm | conv: the other two sharded BASELINE.json configs on 1/2/4/8 GPUs.

dgemm (configs[4], float64 twin): row-sharded DGEMM n^3 (default 32768) — DMMA kernel per row chunk, chunk j's
      ncclAllGather on a communication stream while chunk j+1 computes; C whole on every rank.  Strong scaling.
conv  (configs[3]): LeNet conv2d forward + backward of both layers (cv1 [B,1,28,28]*[20,1,5,5], cv2 [B,20,12,12]*[50,20,5,5]),
      batch B = 4096 split over the ranks, weights replicated, grad_kernel / grad_bias all-reduced (sum).  Strong scaling.
Both print the bench.py JSON line (value = whole-job GFLOP/s, max-over-ranks device time) with a parity field against
the CPU oracle on sampled rows / images (outside the timed region).
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch
import torch.distributed as dist


def _setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local_rank, dev


def _barrier(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _timed(step, args, world, rank, local_rank, dev, sampler_cls, capi):
    for _ in range(args.warmup):
        step()
    _barrier(world)
    sampler = sampler_cls(local_rank)
    if rank == 0:
        sampler.start()
    l0 = capi.kernel_launch_count()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    _barrier(world)
    t0.record()
    for _ in range(args.steps):
        step()
    t1.record()
    _barrier(world)
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    return float(ms.item()) / args.steps, clocks, capi.kernel_launch_count() - l0


def run_dgemm(args, emit, sampler_cls, peaks_fn, log, verify_fn):
    import arraymancer_b200 as am
    from arraymancer_b200 import _capi
    from arraymancer_b200 import distributed as D
    from tools.bench_configs import kostya
    world, rank, local_rank, dev = _setup()
    n = args.n
    chunks = args.chunks if world > 1 else 1
    mc = D.chunk_rows(n, world, chunks)
    rows_local = mc * chunks
    B = kostya(n, dev)                                                  # benchmarks/kostya_matmul.nim generator (A = B)

    def gen_A(r):
        idx = np.concatenate([np.arange((j * world + r) * mc, (j * world + r + 1) * mc) for j in range(chunks)])
        return kostya(n, dev, rows=idx)
    A_local = gen_A(rank)
    sg = D.RowShardedGemm(n, n, n, torch.float64, dev, chunks=chunks)
    ev = []

    def step():
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        sg(A_local, B)
        e1.record()
        ev.append((e0, e1))
    ms, clocks, launches = _timed(step, args, world, rank, local_rank, dev, sampler_cls, _capi)
    flops = 2.0 * n ** 3
    verify = None
    if not args.no_verify:
        verify = verify_fn(am, torch, dist, dev, world, rank, n, rows_local, mc, chunks, B, sg.C, torch.float64, gen_A, 1e-13)
        if world > 1:
            dist.barrier()
    # e2e: per-rank host-buffer entry (pinned): A rows + all of B up, C rows down
    e2e = None
    if not args.no_e2e:
        hA = torch.empty((rows_local, n), dtype=torch.float64, pin_memory=True); hA.copy_(A_local)
        hB = torch.empty((n, n), dtype=torch.float64, pin_memory=True); hB.copy_(B)
        hC = torch.empty((rows_local, n), dtype=torch.float64, pin_memory=True)
        del A_local, B, sg
        torch.cuda.empty_cache()
        lib = _capi.lib()

        def e2e_step():
            _capi.check(lib.am_host_gemm_strided_f64(rows_local, n, n, 1.0, hA.data_ptr(), n, 1, hB.data_ptr(), n, 1, 0.0,
                                                     hC.data_ptr(), n, 1))
        e2e_step()
        _barrier(world)
        w0 = time.perf_counter()
        for _ in range(max(1, min(args.e2e_steps, 3))):
            e2e_step()
        _barrier(world)
        e_ms = 1e3 * (time.perf_counter() - w0) / max(1, min(args.e2e_steps, 3))
        te = torch.tensor([e_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": flops / (float(te.item()) * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": float(te.item()),
               "h2d_bytes_per_step": 8 * (rows_local * n + n * n), "d2h_bytes_per_step": 8 * rows_local * n, "bytes_are": "per rank",
               "note": "am_host_gemm_strided_f64 per rank on pinned host buffers (row chunks: H2D / DMMA GEMM / D2H overlapped)"}
    if rank == 0:
        peaks, src = peaks_fn()
        try:
            dmma_peak = _capi.microbench(4)
        except Exception:  # noqa: BLE001
            dmma_peak = 37.0
        # kernel time: the step is the DMMA launches (+ all-gather overlap); use step time per rank share
        ach = flops / world / (ms * 1e-3) / 1e12
        line = {"metric": "dgemm_gflops", "value": flops / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"row-sharded DGEMM {n}x{n}x{n}, DMMA (BASELINE configs[4], float64)", "M": n, "N": n, "K": n,
                           "parallelism": f"rows of A block-cyclic over {world} GPU(s), B replicated, C all-gathered over NCCL "
                                          f"({chunks} chunk(s)/rank, chunk j's all-gather overlaps chunk j+1's GEMM)",
                           "l2": "operands 8 GiB each >> 126 MB L2 (no flush needed)"},
                "roofline": {"bound": "tensor", "achieved": ach, "peak": dmma_peak, "unit": "TFLOP/s", "frac": ach / dmma_peak,
                             "traffic": None, "kernel": "contract_dmma_kernel", "peak_source": "am_microbench 4 (DMMA m8n8k4), this run",
                             "note": "achieved = per-GPU share of the step (includes the exposed part of the all-gather)"},
                "clocks": clocks, "gpu_launches": int(launches)}
        if e2e:
            line["e2e"] = e2e
        if verify:
            line["verify"] = verify
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_conv(args, emit, sampler_cls, peaks_fn, log):
    import arraymancer_b200 as am
    from arraymancer_b200 import _capi
    from arraymancer_b200 import distributed as D
    world, rank, local_rank, dev = _setup()
    NB = args.batch
    lo, hi = D.shard_batch(NB, world, rank)
    nb = hi - lo
    g = torch.Generator(device=dev); g.manual_seed(2024)                # same stream on every rank: take this rank's images
    X1 = torch.rand((NB, 1, 28, 28), device=dev, generator=g)[lo:hi].contiguous()
    W1 = torch.randn((20, 1, 5, 5), device=dev, generator=g) * float(np.sqrt(2.0 / 25)); B1 = torch.zeros((20, 1, 1), device=dev)
    X2 = torch.rand((NB, 20, 12, 12), device=dev, generator=g)[lo:hi].contiguous()
    W2 = torch.randn((50, 20, 5, 5), device=dev, generator=g) * float(np.sqrt(2.0 / 500)); B2 = torch.rand((50, 1, 1), device=dev, generator=g)
    G1 = (torch.rand((NB, 20, 24, 24), device=dev, generator=g) * 2 - 1)[lo:hi].contiguous()
    G2 = (torch.rand((NB, 50, 8, 8), device=dev, generator=g) * 2 - 1)[lo:hi].contiguous()
    res = {}
    # L2 rule: the per-rank tensors shrink with the rank count (301 MB at 1 GPU, 38 MB at 8) — the step cycles through R
    # identical copies of its inputs so that consecutive steps never find their operands in the 126 MB L2
    per_rank_bytes = nb * (784 + 11520 + 2880 + 3200) * 4
    R = max(1, min(16, -(-300_000_000 // per_rank_bytes)))
    sets = [(X1, X2, G1, G2)] + [tuple(t.clone() for t in (X1, X2, G1, G2)) for _ in range(R - 1)]
    counter = [0]

    def body(x1, x2, g1, g2):
        return {"o1": D.conv2d_batch_sharded(x1, W1, B1), "o2": D.conv2d_batch_sharded(x2, W2, B2),
                "b1": D.conv2d_backward_batch_sharded(x1, W1, B1, (0, 0), (1, 1), (1, 1), g1),
                "b2": D.conv2d_backward_batch_sharded(x2, W2, B2, (0, 0), (1, 1), (1, 1), g2)}

    def step():
        res.update(body(*sets[counter[0] % R]))
        counter[0] += 1
    # The step is ~18 short launches + 4 all-reduces: at 512 images per GPU it is bound by the host's launch rate, not by the
    # kernels.  Capture it once per input copy in a CUDA graph (kernels of this library + the NCCL all-reduces) and replay.
    step(); step()
    torch.cuda.synchronize()
    l0 = _capi.kernel_launch_count(); step(); launches_per_step = _capi.kernel_launch_count() - l0
    graphs, graph_note = None, "eager launches"
    # multi-rank: opt-in (--graph) — replay works (2 GPUs: 0.512 vs 0.522 ms eager) but a process that holds CUDA graphs with
    # captured NCCL kernels hung in its teardown on the 2-GPU box; single GPU: on by default (0.88 vs 1.05 ms eager)
    use_graph = (not getattr(args, "no_graph", False)) and (world == 1 or getattr(args, "graph", False))
    if use_graph:
        try:
            _barrier(world)
            pool = torch.cuda.graph_pool_handle()
            graphs = []
            for i in range(R):
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_, pool=pool, capture_error_mode="relaxed"):
                    out_i = body(*sets[i])
                graphs.append((g_, out_i))
            graph_note = f"{R} CUDA graph(s) (one per input copy), each = the whole step incl. the NCCL all-reduces, replayed"
        except Exception as e:  # noqa: BLE001
            graphs = None
            graph_note = f"eager launches (graph capture failed: {type(e).__name__}: {str(e).splitlines()[0]})"
            log(graph_note)
            for _ in range(3):                                   # a failed capture leaves a sticky error in the library's runtime:
                try:                                             # flush it with collective-free work (the peers are not in this branch)
                    torch.cuda.synchronize(); D.conv2d_batch_sharded(sets[0][0], W1, B1); torch.cuda.synchronize()
                    break
                except Exception:  # noqa: BLE001
                    pass
        flag = torch.tensor([0 if graphs else 1], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)          # all ranks replay, or none does
        if int(flag.item()):
            graphs = None
    if graphs:
        def step():  # noqa: F811
            g_, out_i = graphs[counter[0] % R]
            counter[0] += 1
            g_.replay()
            res.update(out_i)
    ms, clocks, launches = _timed(step, args, world, rank, local_rank, dev, sampler_cls, _capi)
    if graphs:
        launches = launches_per_step * args.steps             # replays do not pass through the library's launch counter
    res = {k: (tuple(t.clone() if t is not None else None for t in v) if isinstance(v, tuple) else v.clone()) for k, v in res.items()}
    graphs = None
    del sets[1:]
    f1 = 2.0 * NB * 20 * 24 * 24 * 25
    f2 = 2.0 * NB * 50 * 8 * 8 * 500
    flops = 3 * (f1 + f2) + NB * (20 * 576 + 50 * 64)                   # fwd + dgrad + wgrad (+ bias sums)
    bytes_ = 4 * ((NB * 784 + 500 + 20 + NB * 11520) + (NB * 2880 + 25000 + 50 + NB * 3200)          # forward
                  + (NB * 784 * 2 + 500 * 2 + 20 + NB * 11520) + (NB * 2880 * 2 + 25000 * 2 + 50 + NB * 3200))
    # ---- parity: every rank checks 8 of its images against the oracle; the all-reduced weight gradients are checked on
    #      rank 0 against the oracle run over the WHOLE batch (serial image order of the reference)
    parity = None
    from oracle import laser_oracle as orc
    orc.build()
    worst = 0.0
    sel = np.linspace(0, nb - 1, num=min(8, nb)).astype(int)
    for (X, W, Bv, G, o, b) in ((X1, W1, B1, G1, res["o1"], res["b1"]), (X2, W2, B2, G2, res["o2"], res["b2"])):
        xs, ws, bs, gs = X[sel].cpu().numpy(), W.cpu().numpy(), Bv.cpu().numpy(), G[sel].cpu().numpy()
        wf = orc.conv2d(xs, ws, bs)
        wgi, _, _ = orc.conv2d_backward(xs, ws, gs)
        rf = np.linalg.norm(o[sel].cpu().numpy().astype(np.float64) - wf) / np.linalg.norm(wf.astype(np.float64))
        rg = np.linalg.norm(b[0][sel].cpu().numpy().astype(np.float64) - wgi) / np.linalg.norm(wgi.astype(np.float64))
        worst = max(worst, float(rf), float(rg))
    tw = torch.tensor([worst], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    gw_rel = None
    if rank == 0:
        g = torch.Generator(device=dev); g.manual_seed(2024)
        X1f = torch.rand((NB, 1, 28, 28), device=dev, generator=g)
        torch.randn((20, 1, 5, 5), device=dev, generator=g)
        X2f = torch.rand((NB, 20, 12, 12), device=dev, generator=g)
        torch.randn((50, 20, 5, 5), device=dev, generator=g); torch.rand((50, 1, 1), device=dev, generator=g)
        G1f = torch.rand((NB, 20, 24, 24), device=dev, generator=g) * 2 - 1
        G2f = torch.rand((NB, 50, 8, 8), device=dev, generator=g) * 2 - 1
        gw_rel = {}
        for name, Xf, W, Gf, b in (("cv1", X1f, W1, G1f, res["b1"]), ("cv2", X2f, W2, G2f, res["b2"])):
            _, wgw, wgb = orc.conv2d_backward(Xf.cpu().numpy(), W.cpu().numpy(), Gf.cpu().numpy())
            gw_rel[name + "_grad_kernel"] = float(np.linalg.norm(b[1].cpu().numpy().astype(np.float64) - wgw) / np.linalg.norm(wgw.astype(np.float64)))
            gw_rel[name + "_grad_bias"] = float(np.linalg.norm(b[2].cpu().numpy().astype(np.float64) - wgb) / np.linalg.norm(wgb.astype(np.float64)))
        parity = {"ok": bool(float(tw.item()) <= 5e-6 and max(gw_rel.values()) <= 1e-4), "kind": "rel_fro",
                  "value": max(float(tw.item()), max(gw_rel.values())), "fwd_and_grad_input_max": float(tw.item()),
                  "weight_gradients": gw_rel, "tol": {"forward / grad_input": 5e-6, "grad_kernel / grad_bias": 1e-4},
                  "against": "oracle.conv2d / conv2d_backward (restated im2colgemm_conv2d(+_gradient))",
                  "sample": f"8 images per rank (forward, grad_input); all-reduced grad_kernel / grad_bias vs the oracle over all {NB} images"}
    if world > 1:
        dist.barrier()
    # ---- e2e: images + grad_outputs from pinned host memory, gradients back to the host, every step
    e2e = None
    if not args.no_e2e:
        hX1 = torch.empty(X1.shape, pin_memory=True).copy_(X1); hX2 = torch.empty(X2.shape, pin_memory=True).copy_(X2)
        hG1 = torch.empty(G1.shape, pin_memory=True).copy_(G1); hG2 = torch.empty(G2.shape, pin_memory=True).copy_(G2)
        hO1 = torch.empty(G1.shape, pin_memory=True); hO2 = torch.empty(G2.shape, pin_memory=True)
        hI1 = torch.empty(X1.shape, pin_memory=True); hI2 = torch.empty(X2.shape, pin_memory=True)
        hW = [torch.empty(t.shape, pin_memory=True) for t in (W1, B1, W2, B2)]

        def e2e_step():
            x1 = hX1.to(dev, non_blocking=True); x2 = hX2.to(dev, non_blocking=True)
            g1 = hG1.to(dev, non_blocking=True); g2 = hG2.to(dev, non_blocking=True)
            o1 = D.conv2d_batch_sharded(x1, W1, B1); o2 = D.conv2d_batch_sharded(x2, W2, B2)
            b1 = D.conv2d_backward_batch_sharded(x1, W1, B1, (0, 0), (1, 1), (1, 1), g1)
            b2 = D.conv2d_backward_batch_sharded(x2, W2, B2, (0, 0), (1, 1), (1, 1), g2)
            hO1.copy_(o1, non_blocking=True); hO2.copy_(o2, non_blocking=True)
            hI1.copy_(b1[0], non_blocking=True); hI2.copy_(b2[0], non_blocking=True)
            for h, t in zip(hW, (b1[1], b1[2], b2[1], b2[2])):
                h.copy_(t, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e2e_step()
        _barrier(world)
        w0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        _barrier(world)
        e_ms = 1e3 * (time.perf_counter() - w0) / args.e2e_steps
        te = torch.tensor([e_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        per_img_in = 4 * (784 + 2880 + 11520 + 3200)
        per_img_out = 4 * (11520 + 3200 + 784 + 2880)
        e2e = {"value": flops / (float(te.item()) * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": float(te.item()),
               "h2d_bytes_per_step": nb * per_img_in, "d2h_bytes_per_step": nb * per_img_out + 4 * (500 + 20 + 25000 + 50),
               "bytes_are": "per rank", "note": "images and grad_outputs from pinned host memory, outputs / grad_inputs / weight gradients back to the host"}
    if rank == 0:
        peaks, src = peaks_fn()
        hbm = float(peaks.get("hbm_gbs", 6454.3))
        ach = bytes_ / world / (ms * 1e-3) / 1e9
        emit({"metric": "conv2d_fwd_bwd_gflops", "value": flops / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "n_gpus": world,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
              "vs_baseline": None, "dtype": "f32", "data": "synthetic",
              "config": {"workload": f"LeNet conv2d cv1+cv2 forward+backward, batch {NB} (BASELINE configs[3])", "batch": NB,
                         "parallelism": f"batch split over {world} GPU(s), weights replicated, grad_kernel/grad_bias all-reduced (NCCL)",
                         "launch": graph_note,
                         "l2": f"per-rank tensors {per_rank_bytes / 1e6:.0f} MB; the step cycles through {R} copies of its inputs "
                               f"({R * per_rank_bytes / 1e6:.0f} MB > 2 x 126 MB L2), so no step finds its operands in L2"},
              "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
                           "kernel": "conv fwd/dgrad/wgrad kernels of both layers", "note": "algorithmic bytes of the four calls / step time, per GPU"},
              "clocks": clocks, "gpu_launches": int(launches), "parity": parity, **({"e2e": e2e} if e2e else {})})
    if world > 1:
        dist.destroy_process_group()
