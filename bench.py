#!/usr/bin/env python
"""bench.py — headline benchmark of the dense-contraction hot path (BASELINE.json configs[4]):
row-sharded float32 SGEMM 32768^2 (3xTF32 on tcgen05, fp32 accuracy) at 1/2/4/8 B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                      # the reference's CPU algorithm (oracle port)

One step = one full C = A*B: every rank splits/packs its A rows and the replicated B, runs the
tcgen05 GEMM on its block-cyclic row chunks and all-gathers C over NCCL (chunk j's all-gather
overlaps chunk j+1's GEMM).  Inputs are resident in HBM before the timed region; `e2e` repeats the
measurement through the host-buffer C-ABI entry (pinned host memory, H2D + D2H inside the region).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sgemm_gflops"
UNIT = "GFLOP/s"


_REAL_STDOUT = None


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.stop_flag, self.t = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.15)

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.t:
            self.t.join(timeout=6)
        sm = sorted(int(float(s[0])) for s in self.samples if s and s[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for n, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        smax = int(float(self.samples[0][1])) if self.samples else None
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(self.samples)}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for this path (restated laser gemm_strided,
    oracle/ — the Nim original cannot be built here), all host threads, on a bounded row sample of the
    same 32768^2 workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and may use the whole host
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("OMP_NUM_THREADS", "") == "1":
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    import numpy as np
    from oracle import laser_oracle as orc
    orc.build()
    n = args.n
    rows = args.cpu_rows
    rng = np.random.default_rng(1235)
    B = rng.random((n, n), dtype=np.float32) * 2 - 1
    A = np.random.default_rng(1234).random((rows, n), dtype=np.float32) * 2 - 1
    C = np.empty((rows, n), dtype=np.float32)
    threads = orc.max_threads()
    for _ in range(max(1, min(args.warmup, 1))):
        orc.gemm_strided(1.0, A[:64], B, 0.0, C[:64])
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        orc.gemm_strided(1.0, A, B, 0.0, C)
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * sum(ts) / len(ts)
    val = 2.0 * rows * n * n / (ms * 1e-3) / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"row-sharded SGEMM {n}x{n}x{n} (BASELINE configs[4]); CPU arm runs a {rows}-row sample",
                       "M": n, "N": n, "K": n},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"rows 0..{rows - 1} of A against the full {n}x{n} B (full N and K), restated laser "
                                       f"gemm_strided f32 AVX2+FMA 6x16 micro-kernel, OpenMP {threads} threads"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=32768, help="square size (BASELINE configs[4]: 32768)")
    ap.add_argument("--chunks", type=int, default=4, help="row chunks per rank for compute/all-gather overlap (N>1)")
    ap.add_argument("--cpu-rows", type=int, default=768, help="rows of A in the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--comm", default="fused", choices=["fused", "nccl"],
                    help="N>1: 'fused' = the GEMM epilogue stores C to every GPU's copy over NVLink (symmetric memory); "
                         "'nccl' = per-chunk ncclAllGather on a second stream")
    ap.add_argument("--verify", action="store_true", help="check the gathered C against a local recomputation of sampled rows of every rank")
    args = ap.parse_args()
    # NCCL prints a version banner on stdout at communicator creation: keep stdout clean for the ONE JSON line
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import arraymancer_b200 as am
    from arraymancer_b200 import _capi
    from arraymancer_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback on the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.n
    chunks = args.chunks if world > 1 else 1
    mc = D.chunk_rows(n, world, chunks)
    rows_local = mc * chunks

    # ---- synthetic inputs, resident in HBM (A: this rank's block-cyclic rows; B replicated, same seed)
    gB = torch.Generator(device=dev); gB.manual_seed(1235)
    B = torch.rand((n, n), device=dev, dtype=torch.float32, generator=gB) * 2 - 1
    gA = torch.Generator(device=dev); gA.manual_seed(1234 + 7919 * rank)
    A_local = torch.rand((rows_local, n), device=dev, dtype=torch.float32, generator=gA) * 2 - 1
    comm = args.comm if world > 1 else "none"
    symC = None
    if comm == "fused":
        try:
            symC = D.SymmetricResult((n, n), torch.float32, dev)
            C = symC.C
        except Exception as e:  # no peer mapping on this box: fall back to the NCCL collective, and say so
            print(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); using --comm nccl", file=sys.stderr)
            comm = "nccl"
            flag = torch.tensor([1], device=dev)
        else:
            flag = torch.tensor([0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)                   # all ranks take the same path
        if int(flag.item()) and comm == "fused":
            comm, symC = "nccl", None
    if symC is None:
        C = torch.empty((n, n), device=dev, dtype=torch.float32)      # full result on every rank
    if comm == "fused" and args.chunks == 4:
        chunks = 1                                                    # nothing to overlap: one launch per rank and step
        mc = D.chunk_rows(n, world, chunks)
        rows_local = mc * chunks
        A_local = A_local[:rows_local]
    comm_stream = torch.cuda.Stream(device=dev) if comm == "nccl" else None

    pB = am.PackedF32(B, "b")
    pA = am.PackedF32(A_local[:mc], "a")
    ev_k0, ev_k1 = [], []
    launches0 = 0

    def step(record_kernel=False):
        works = []
        pB.repack(B)                                          # split/pack B (hi/lo tf32 planes) — part of the job
        for j in range(chunks):
            lo = (j * world + rank) * mc
            mine = C[lo:lo + mc]
            pA.repack(A_local[j * mc:(j + 1) * mc])
            if record_kernel:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            if comm == "fused":
                am.gemm_packed_bcast(1.0, pA, pB, mine, symC.peer_ptrs(mine), symC.rank)   # mainloop + stores to all GPUs
            else:
                am.gemm_packed(1.0, pA, pB, 0.0, mine)        # tcgen05 3xTF32 mainloop
            if record_kernel:
                e1.record(); ev_k0.append(e0); ev_k1.append(e1)
            if comm == "nccl":
                span = C[j * world * mc:(j + 1) * world * mc]
                ready = torch.cuda.Event(); ready.record()
                with torch.cuda.stream(comm_stream):
                    comm_stream.wait_event(ready)
                    works.append(dist.all_gather_into_tensor(span, mine, async_op=True))
        for w in works:
            w.wait()
        if comm == "nccl":
            torch.cuda.current_stream().wait_stream(comm_stream)
        if comm == "fused":
            symC.barrier()                                    # every rank's tiles have landed in every copy of C

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _capi.kernel_launch_count()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        step(record_kernel=True)
    t1.record()
    barrier()
    total_ms = t0.elapsed_time(t1)
    launches = _capi.kernel_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    kern_ms = sum(a.elapsed_time(b) for a, b in zip(ev_k0, ev_k1)) / max(1, len(ev_k0))
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    flops_step = 2.0 * n * n * n
    value = flops_step / (ms_per_step * 1e-3) / 1e9

    verify = None
    if args.verify:
        # every rank recomputes 64 rows of every rank's first chunk from that rank's seed and compares with the gathered C
        worst = 0.0
        for r in range(world):
            gr = torch.Generator(device=dev); gr.manual_seed(1234 + 7919 * r)
            Ar = torch.rand((rows_local, n), device=dev, dtype=torch.float32, generator=gr) * 2 - 1
            ref = torch.empty((64, n), device=dev, dtype=torch.float32)
            am.gemm_strided(1, Ar[:64], B, 0, ref)
            lo = r * mc
            got = C[lo:lo + 64]
            worst = max(worst, float(((got - ref).double().norm() / ref.double().norm()).item()))
            del Ar
        tv = torch.tensor([worst], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        verify = {"max_rel_fro_vs_single_gpu_rows": float(tv.item()), "ok": bool(tv.item() <= 1e-6)}

    # ---- e2e: same job through the host-buffer C-ABI entry (pinned host memory, copies inside the region)
    e2e = None
    if not args.no_e2e:
        del pA, pB
        hA = torch.empty((rows_local, n), dtype=torch.float32, pin_memory=True)
        hB = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
        hC = torch.empty((rows_local, n), dtype=torch.float32, pin_memory=True)
        hA.copy_(A_local); hB.copy_(B)
        del A_local, B, C
        torch.cuda.empty_cache()
        lib = _capi.lib()
        def e2e_step():
            _capi.check(lib.am_host_gemm_strided_f32(rows_local, n, n, 1.0, hA.data_ptr(), n, 1, hB.data_ptr(), n, 1,
                                                     0.0, hC.data_ptr(), n, 1))
        e2e_step()
        barrier()
        w0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        e_ms = 1e3 * (time.perf_counter() - w0) / args.e2e_steps
        te = torch.tensor([e_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_ms = float(te.item())
        e2e = {"value": flops_step / (e_ms * 1e-3) / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": 4 * (rows_local * n + n * n), "d2h_bytes_per_step": 4 * rows_local * n,
               "ms_per_step": e_ms,
               "note": "am_host_gemm_strided_f32 per rank on pinned host buffers: H2D of the rank's A rows and of B, "
                       "tcgen05 GEMM, D2H of the rank's C rows; wall clock, max over ranks"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, src = _peaks()
    # 3xTF32: tf32 rate = bf16/2, three MMAs per product.  Kernel timed inside a long step -> sustained figure.
    peak_tf = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) / 6.0
    kern_flops = 2.0 * mc * n * n                      # algorithmic FLOPs of ONE mainloop launch (one row chunk)
    achieved_tf = kern_flops / (kern_ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("gemm_tf32x3_kernel", {}).get(f"n{n}_g{world}")
        except Exception:
            traffic = None
    # the library's own tcgen05 kind::tf32 micro-benchmark (am_microbench 5, profiles/) gives a second denominator
    own_peak = None
    try:
        for ln in open(os.path.join(ROOT, "profiles", "r01_bringup_final.jsonl")):
            d = json.loads(ln)
            if d.get("exp") == "peaks":
                own_peak = float(d["result"]["umma_tf32_1cta"]) / 3.0
    except Exception:
        own_peak = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"row-sharded SGEMM {n}x{n}x{n}, 3xTF32 on tcgen05 (BASELINE configs[4])",
                   "M": n, "N": n, "K": n, "parallelism": f"rows of A block-cyclic over {world} GPU(s), B replicated, "
                   + {"fused": "C tiles stored to every GPU's copy by the GEMM epilogue (symmetric memory over NVLink, no separate collective)",
                      "nccl": f"C all-gathered over NCCL ({chunks} chunk(s)/rank, overlapped)",
                      "none": "single GPU"}[comm],
                   "comm": comm, "chunks": chunks,
                   "l2": "operands 4 GiB each >> 126 MB L2 (no flush needed)",
                   "timed": "split/pack of A rows and B + tcgen05 mainloop + " +
                            ("stores of C to all GPUs + cross-GPU barrier" if comm == "fused" else "all-gather of C")},
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf, "traffic": traffic,
                     "kernel": "gemm_tf32x3_kernel<2>", "kernel_ms": kern_ms, "flops_per_launch": kern_flops,
                     "peak_source": f"{src}: bf16_tflops_sustained / 2 (tf32 rate) / 3 (three tf32 MMAs per fp32 product)",
                     "own_umma_tf32_peak_div3": own_peak,
                     "frac_of_own_umma_peak": (achieved_tf / own_peak) if own_peak else None},
        "clocks": clocks, "gpu_launches": int(launches),
    }
    if e2e:
        line["e2e"] = e2e
    if verify:
        line["verify"] = verify
    if world == 1 and not args.no_cpu:
        # CPU baseline: restated laser gemm_strided on the host cores, bounded row sample of the same problem
        from oracle import laser_oracle as orc
        orc.build()
        rows = args.cpu_rows
        hb = hB.numpy() if not args.no_e2e else (np.random.default_rng(1235).random((n, n), dtype=np.float32) * 2 - 1)
        ha = (hA.numpy() if not args.no_e2e else np.random.default_rng(1234).random((rows, n), dtype=np.float32))[:rows]
        hc = np.empty((rows, n), dtype=np.float32)
        orc.gemm_strided(1.0, ha[:32], hb, 0.0, hc[:32])
        c0 = time.perf_counter()
        orc.gemm_strided(1.0, ha, hb, 0.0, hc)
        cs = time.perf_counter() - c0
        line["cpu_baseline"] = {"value": 2.0 * rows * n * n / cs / 1e9, "unit": UNIT, "cores": orc.max_threads(),
                                "kind": "port", "seconds": cs,
                                "sample": f"rows 0..{rows - 1} of A against the full {n}x{n} B (full N and K); restated laser "
                                          "gemm_strided f32 (AVX2+FMA 6x16 micro-kernel, OpenMP), not a Nim build"}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
