#!/usr/bin/env python
"""bench.py — benchmark of the dense-contraction hot path.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (default workload: sgemm)
  python bench.py --impl reference ...                      # the reference's CPU algorithm (oracle port)
  python bench.py --workload dgemm|conv --gpus N ...        # the other sharded BASELINE configs

Headline (`--workload sgemm`, BASELINE.json configs[4]): row-sharded float32 SGEMM 32768^2 (3xTF32 on tcgen05, fp32
accuracy) at 1/2/4/8 B200.  One step = one full C = A*B: every rank splits/packs its A rows and B, runs the tcgen05
GEMM on its rows and makes C whole on every rank (fused: the GEMM epilogue stores to every GPU's copy of C over
NVLink; nccl: chunked all-gather on a second stream).  Inputs are resident in HBM before the timed region.
`e2e` repeats the job from pinned HOST buffers (H2D + D2H inside the region) through the host-buffer entries.
At N = 1 the line also carries:
  * `verify`  — sampled rows of the timed C against the CPU ORACLE (restated laser gemm_strided) at full K;
  * `configs` — every other BASELINE config (C1 int64/int32, C2 float64 8192^2, C3 float32 16384^2 in six layouts,
                 C4 LeNet conv fwd/bwd at batch 4096, C5 float64 32768^2, skinny DRAM-bound products), each with its own
                 time, roofline fraction, clock sample and oracle parity (tools/bench_configs.py);
  * `cpu_baseline` — the oracle port on the box's host cores, bounded sample, median of 5, both ISA variants.
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

UNIT = "GFLOP/s"
METRICS = {"sgemm": "sgemm_gflops", "dgemm": "dgemm_gflops", "conv": "conv2d_fwd_bwd_gflops"}

_REAL_STDOUT = None


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


def _log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


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
            time.sleep(0.1)

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


# ================================================================================================ CPU arm
def _host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _cpu_gemm_sample(orc, n, rows, reps, variants, log=_log):
    """Times the restated laser gemm_strided (f32) on `rows` rows of A against the full n x n B.  Returns
    {variant_name: {"gflops": median, "seconds": [...]} }.  rows = 192 * threads => one mc = 192 ic tile per thread;
    the jr task loop (gemm.nim:82, restated in oracle/laser_gemm.hpp) balances the rest."""
    import numpy as np
    B = np.random.default_rng(1235).random((n, n), dtype=np.float32) * 2 - 1
    A = np.random.default_rng(1234).random((rows, n), dtype=np.float32) * 2 - 1
    C = np.empty((rows, n), dtype=np.float32)
    out = {}
    for name, lib_mod, var in variants:
        lib_mod.gemm_strided(1.0, A[:64], B, 0.0, C[:64], variant=var)          # warm: thread pool, page faults
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            lib_mod.gemm_strided(1.0, A, B, 0.0, C, variant=var)
            ts.append(time.perf_counter() - t0)
        ts.sort()
        med = ts[len(ts) // 2]
        out[name] = {"gflops": 2.0 * rows * n * n / med / 1e9, "seconds_median": med, "reps": reps}
        log(f"cpu {name}: rows={rows} median {med:.3f} s -> {out[name]['gflops']:.1f} GFLOP/s")
    return out


def _cpu_variants(orc):
    """(name, module, variant) list: the default-ISA build (AVX2+FMA 6x16, what the published numbers ran) and the
    -d:avx512 tile shapes (14x32), the latter compiled -march=native on this host when it has AVX-512."""
    v = [("default_avx2_fma_6x16", orc, orc.DEFAULT_BUILD)]
    try:
        nat = orc.native()
        if nat is not None:
            v.append(("avx512_tiles_14x32_march_native", nat, orc.AVX512))
    except Exception as e:  # noqa: BLE001
        _log(f"native oracle build unavailable: {e}")
    return v


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for this path (restated laser gemm_strided, oracle/ — the
    Nim original cannot be built here), all host threads, each step a bounded row sample of the same 32768^2 job."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = _host_threads()
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and may use the whole host
    if os.environ.get("OMP_NUM_THREADS", "") in ("", "1") and threads > 1:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    import numpy as np
    from oracle import laser_oracle as orc
    orc.build()
    threads = orc.max_threads()
    n = args.n
    if args.workload == "dgemm":
        dt, npdt, mc = "f64", np.float64, 96
    else:
        dt, npdt, mc = "f32", np.float32, 192
    rows = args.cpu_rows or min(n, mc * threads)
    B = (np.random.default_rng(1235).random((n, n)) * 2 - 1).astype(npdt)
    A = (np.random.default_rng(1234).random((rows, n)) * 2 - 1).astype(npdt)
    C = np.empty((rows, n), dtype=npdt)
    orc.gemm_strided(1.0, A[:64], B, 0.0, C[:64])
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        orc.gemm_strided(1.0, A, B, 0.0, C)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    ms = 1e3 * ts[len(ts) // 2]
    val = 2.0 * rows * n * n / (ms * 1e-3) / 1e9
    other = None
    if dt == "f32" and not args.no_cpu_avx512:
        try:
            vs = [v for v in _cpu_variants(orc) if v[0].startswith("avx512")]
            if vs:
                other = _cpu_gemm_sample(orc, n, rows, min(3, max(1, args.steps)), vs)
        except Exception as e:  # noqa: BLE001
            _log(f"avx512 variant skipped: {e}")
    metric = METRICS["dgemm" if dt == "f64" else "sgemm"]
    sample = (f"rows 0..{rows - 1} of A (= {mc} x {threads} threads: one mc tile per thread, jr task loop as in gemm.nim:82) "
              f"against the full {n}x{n} B (full N and K; the whole pack_B of every kc panel is charged to the sample); "
              f"restated laser gemm_strided {dt} default-ISA micro-kernel ({'6x16' if dt == 'f32' else '6x8'} AVX2+FMA), "
              f"OpenMP {threads} threads; median of {args.steps} steps")
    line = {"impl": "reference", "metric": metric, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": dt, "data": "synthetic",
            "config": {"workload": f"row-sharded {'DGEMM' if dt == 'f64' else 'SGEMM'} {n}x{n}x{n} (BASELINE configs[4]); "
                                   f"CPU arm runs a {rows}-row sample per step", "M": n, "N": n, "K": n, "sample_rows": rows},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "variants": other},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# ================================================================================================ helpers
def _verify_vs_oracle(am, torch, dist, dev, world, rank, n, rows_local, mc, chunks, B, C, dtype, seed_fn, tol):
    """Sampled rows of the timed C (8 per rank's block) against oracle.gemm_strided at full K — on rank 0."""
    import numpy as np
    if rank != 0:
        return None
    from oracle import laser_oracle as orc
    orc.build()
    rng = np.random.default_rng(99)
    a_rows, got_rows = [], []
    per = max(1, 64 // world)
    for r in range(world):
        A_r = seed_fn(r)                                     # rank r's local rows, regenerated from its seed
        loc = np.sort(rng.choice(rows_local, size=per, replace=False))
        for li in loc:
            j, off = divmod(int(li), mc)
            grow = (j * world + r) * mc + off                # block-cyclic global row of local row li
            a_rows.append(A_r[int(li)].cpu().numpy())
            got_rows.append(C[grow].cpu().numpy())
        del A_r
    a = np.stack(a_rows); got = np.stack(got_rows)
    b = B.cpu().numpy()
    want = np.zeros_like(got)
    t0 = time.perf_counter()
    orc.gemm_strided(1.0, a, b, 0.0, want)
    rel = float(np.linalg.norm(got.astype(np.float64) - want.astype(np.float64)) / np.linalg.norm(want.astype(np.float64)))
    return {"ok": bool(rel <= tol and np.isfinite(got).all()), "kind": "rel_fro", "value": rel, "tol": tol,
            "against": "oracle.gemm_strided (restated laser gemm_strided, CPU) at full K",
            "sample": f"{len(a_rows)} rows of the gathered C ({per} from every rank's block), all {n} columns, K = {n}",
            "oracle_seconds": time.perf_counter() - t0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sgemm", choices=["sgemm", "dgemm", "conv"])
    ap.add_argument("--size", dest="n", type=int, default=32768, help="square size (BASELINE configs[4]: 32768)")
    ap.add_argument("--batch", type=int, default=4096, help="conv workload: total images (BASELINE configs[3]: 4096)")
    ap.add_argument("--chunks", type=int, default=4, help="row chunks per rank for compute/all-gather overlap (N>1)")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of A in the bounded CPU sample (0 = 192 x threads)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-cpu-avx512", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (other BASELINE configs, N=1)")
    ap.add_argument("--configs-only", default="", help="comma list of config prefixes (C1,C2,C3,C4,C5,skinny)")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--comm", default="fused", choices=["fused", "nccl"],
                    help="N>1: 'fused' = the GEMM epilogue stores C to every GPU's copy over NVLink (symmetric memory); "
                         "'nccl' = per-chunk ncclAllGather on a second stream")
    ap.add_argument("--no-graph", action="store_true", help="conv workload: launch eagerly instead of replaying CUDA graphs")
    ap.add_argument("--graph", action="store_true", help="conv workload, N>1: replay CUDA graphs (incl. the NCCL all-reduces) too")
    ap.add_argument("--static-b", action="store_true",
                    help="sgemm: B is a constant (weight-like) operand, packed once outside the timed step")
    args = ap.parse_args()
    # NCCL prints a version banner on stdout at communicator creation: keep stdout clean for the ONE JSON line
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback on the product path)")
    if args.workload == "conv":
        from tools import bench_workloads
        return bench_workloads.run_conv(args, _emit, ClockSampler, _peaks, _log)
    if args.workload == "dgemm":
        from tools import bench_workloads
        return bench_workloads.run_dgemm(args, _emit, ClockSampler, _peaks, _log, _verify_vs_oracle)
    return run_sgemm(args)


def run_sgemm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import arraymancer_b200 as am
    from arraymancer_b200 import _capi
    from arraymancer_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.n
    chunks = args.chunks if world > 1 else 1
    comm = args.comm if world > 1 else "none"
    if comm == "fused" and args.chunks == 4:
        chunks = 1                                                    # nothing to overlap: one launch per rank and step
    mc = D.chunk_rows(n, world, chunks)
    rows_local = mc * chunks

    # ---- synthetic inputs, resident in HBM (A: this rank's block-cyclic rows; B replicated, same seed)
    def gen_A(r):
        g = torch.Generator(device=dev); g.manual_seed(1234 + 7919 * r)
        return torch.rand((rows_local, n), device=dev, dtype=torch.float32, generator=g) * 2 - 1
    gB = torch.Generator(device=dev); gB.manual_seed(1235)
    B = torch.rand((n, n), device=dev, dtype=torch.float32, generator=gB) * 2 - 1
    A_local = gen_A(rank)
    symC = None
    if comm == "fused":
        try:
            symC = D.SymmetricResult((n, n), torch.float32, dev)
            C = symC.C
            flag = torch.tensor([0], device=dev)
        except Exception as e:  # no peer mapping on this box: fall back to the NCCL collective, and say so
            print(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); using --comm nccl", file=sys.stderr)
            flag = torch.tensor([1], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)                   # all ranks take the same path
        if int(flag.item()):
            comm, symC = "nccl", None
    if symC is None:
        C = torch.empty((n, n), device=dev, dtype=torch.float32)      # full result on every rank
    comm_stream = torch.cuda.Stream(device=dev) if comm == "nccl" else None

    pB = am.PackedF32(B, "b")
    pA = am.PackedF32(A_local[:mc], "a")
    ev_k0, ev_k1 = [], []

    def step(record_kernel=False):
        works = []
        if comm == "fused":
            symC.barrier()                                    # peers have finished reading the previous step's C
        if not args.static_b:
            pB.repack(B)                                      # split/pack B (hi/lo tf32 planes) — part of the job
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
    if not args.no_verify:
        verify = _verify_vs_oracle(am, torch, dist, dev, world, rank, n, rows_local, mc, chunks, B, C, torch.float32, gen_A, 5e-6)
        if world > 1:
            dist.barrier()

    # ---- e2e: same job from pinned HOST buffers (copies inside the region)
    e2e = None
    if not args.no_e2e:
        del pA, pB
        hA = torch.empty((rows_local, n), dtype=torch.float32, pin_memory=True)
        hC = torch.empty((rows_local, n), dtype=torch.float32, pin_memory=True)
        hA.copy_(A_local)
        if world == 1:
            hB = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
            hB.copy_(B)
        else:
            # every rank holds (only) 1/g of B in host memory: its part of each of the S K slices (HostShardedGemmF32)
            S = 4
            kc = n // S
            part = kc // world
            hB = torch.empty((n // world, n), dtype=torch.float32, pin_memory=True)
            for c in range(S):
                hB[c * part:(c + 1) * part].copy_(B[c * kc + rank * part:c * kc + (rank + 1) * part])
        C_dev_rows = C[rank * mc:rank * mc + 8].clone() if chunks == 1 else None    # for the e2e self-check below
        del A_local, B, C, symC
        torch.cuda.empty_cache()
        lib = _capi.lib()
        if world == 1:
            def e2e_step():
                _capi.check(lib.am_host_gemm_strided_f32(rows_local, n, n, 1.0, hA.data_ptr(), n, 1, hB.data_ptr(), n, 1,
                                                         0.0, hC.data_ptr(), n, 1))
            note = ("am_host_gemm_strided_f32 on pinned host buffers: K-pipelined H2D of A and B, split/pack, tcgen05 GEMM, "
                    "D2H of C by row chunks; wall clock")
            h2d, d2h = 4 * (rows_local * n + n * n), 4 * rows_local * n
        else:
            hs = D.HostShardedGemmF32(rows_local, n, n, dev, slices=S)
            def e2e_step():
                hs(hA, hB, hC)
            note = ("HostShardedGemmF32: K in 4 slices; per slice every rank uploads its rows of A (pitched copy) and ONLY its 1/g "
                    "part of B's slice into symmetric memory, pulls the other parts from the peers with the copy engines over "
                    "NVLink, splits/packs and accumulates C_local += A_c*B_c on tcgen05 while the next slice uploads; the last "
                    "slice runs by row chunks so C streams back to the host; wall clock, max over ranks")
            h2d, d2h = 4 * (rows_local * n + (n // world) * n), 4 * rows_local * n
        e2e_step()
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
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "bytes_are": "per rank",
               "ms_per_step": e_ms, "steps": args.e2e_steps, "note": note}
        if C_dev_rows is not None:
            got = hC[:8].to(dev)
            e2e["rel_fro_vs_device_resident_rows"] = float(((got - C_dev_rows).double().norm() / C_dev_rows.double().norm()).item())
    else:
        del pA, pB, A_local, B, C
        torch.cuda.empty_cache()

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
    line = {
        "metric": METRICS["sgemm"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"row-sharded SGEMM {n}x{n}x{n}, 3xTF32 on tcgen05 (BASELINE configs[4])",
                   "M": n, "N": n, "K": n, "parallelism": f"rows of A block-cyclic over {world} GPU(s), B replicated, "
                   + {"fused": "C tiles stored to every GPU's copy by the GEMM epilogue (symmetric memory over NVLink, no separate collective)",
                      "nccl": f"C all-gathered over NCCL ({chunks} chunk(s)/rank, overlapped)",
                      "none": "single GPU"}[comm],
                   "comm": comm, "chunks": chunks, "static_b": bool(args.static_b),
                   "l2": "operands 4 GiB each >> 126 MB L2 (no flush needed)",
                   "timed": ("" if args.static_b else "split/pack of B + ") + "split/pack of A rows + tcgen05 mainloop + " +
                            ("stores of C to all GPUs + cross-GPU barriers" if comm == "fused" else
                             "all-gather of C" if comm == "nccl" else "stores of C")},
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf, "traffic": traffic,
                     "kernel": "gemm_tf32x3_kernel<2>", "kernel_ms": kern_ms, "flops_per_launch": kern_flops,
                     "peak_source": f"{src}: bf16_tflops_sustained / 2 (tf32 rate) / 3 (three tf32 MMAs per fp32 product)"},
        "clocks": clocks, "gpu_launches": int(launches),
    }
    if e2e:
        line["e2e"] = e2e
    if verify:
        line["verify"] = verify
    if world == 1 and not args.no_configs:
        from oracle import laser_oracle as orc
        orc.build()
        from tools import bench_configs
        only = [s for s in args.configs_only.split(",") if s] or None
        try:
            recs, pipe = bench_configs.run_configs(am, orc, peaks, ClockSampler, local_rank, only=only, log=_log)
            line["configs"] = recs
            line["pipe_peaks"] = pipe
            if pipe.get("umma_tf32_2cta"):
                line["roofline"]["own_umma_tf32_peak_div3"] = pipe["umma_tf32_2cta"] / 3.0
                line["roofline"]["frac_of_own_umma_peak"] = achieved_tf / (pipe["umma_tf32_2cta"] / 3.0)
        except Exception as e:  # noqa: BLE001 — the headline must still be printed
            import traceback
            traceback.print_exc()
            line["configs_error"] = f"{type(e).__name__}: {e}"
    if world == 1 and not args.no_cpu:
        # CPU baseline: restated laser gemm_strided on the host cores, bounded row sample of the same problem
        from oracle import laser_oracle as orc
        orc.build()
        threads = orc.max_threads()
        rows = args.cpu_rows or min(n, 96 * threads)
        res = _cpu_gemm_sample(orc, n, rows, 5, _cpu_variants(orc) if not args.no_cpu_avx512 else _cpu_variants(orc)[:1])
        d0 = res["default_avx2_fma_6x16"]
        line["cpu_baseline"] = {"value": d0["gflops"], "unit": UNIT, "cores": threads, "kind": "port",
                                "seconds": d0["seconds_median"], "variants": res,
                                "sample": f"rows 0..{rows - 1} of A (96 x {threads} threads) against the full {n}x{n} B (full N "
                                          "and K); restated laser gemm_strided f32 (default-ISA AVX2+FMA 6x16 micro-kernel, "
                                          "OpenMP ic loop + jr task loop), not a Nim build; median of 5"}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
