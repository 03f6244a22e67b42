"""CPU: the parts of the bench.py contract that can be checked without a GPU — the reference arm prints ONE JSON line
with the agreed keys (and only rank 0 speaks under torchrun), and our arm refuses to run without a GPU (no CPU
fallback on the product path)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.pop("RANK", None); e.pop("WORLD_SIZE", None)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_line():
    p = _run(["--impl", "reference", "--size", "512", "--cpu-rows", "64", "--steps", "1", "--warmup", "1", "--gpus", "1"])
    assert p.returncode == 0, p.stderr[-500:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["config"]["sample_rows"] == 64 and "jr task loop" in d["cpu_baseline"]["sample"]


def test_reference_arm_default_sample_scales_with_threads():
    """rows = 192 x threads (one mc tile per thread) unless --cpu-rows is given; OMP team = all host threads."""
    p = _run(["--impl", "reference", "--size", "768", "--steps", "1", "--gpus", "1", "--no-cpu-avx512"])
    assert p.returncode == 0, p.stderr[-500:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    cores = d["cpu_baseline"]["cores"]
    assert cores == len(os.sched_getaffinity(0))
    assert d["config"]["sample_rows"] == min(768, 192 * cores)


def test_device_generator_matches_host_splitmix64():
    import numpy as np
    import torch
    from tools import bench_configs as bc
    from tests.conftest import splitmix64
    z = bc._splitmix64_torch(42, 4096, "cpu").numpy().view(np.uint64)
    assert np.array_equal(z, splitmix64(42, 4096)) and np.array_equal(z, bc.splitmix64_numpy(42, 4096))
    m = bc.gen_matrix_chunked("u100", 37, 53, 42, "cpu", torch.int64).numpy()
    want = ((splitmix64(42, 37 * 53) >> np.uint64(33)) % np.uint64(100)).astype(np.int64).reshape(37, 53)
    assert np.array_equal(m, want)
    f = bc.gen_matrix_chunked("u11", 5, 7, 1234, "cpu", torch.float32).numpy()
    assert f.min() >= -1 and f.max() < 1
    k = bc.kostya(6, "cpu").numpy()
    i, j = np.arange(6.0)[:, None], np.arange(6.0)[None, :]
    assert np.array_equal(k, (1.0 / 36) * (i - j) * (i + j))


def test_reference_arm_other_ranks_stay_silent():
    p = _run(["--impl", "reference", "--size", "512", "--steps", "1", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return                                     # on a GPU box the real bench is exercised by the driver
    p = _run(["--steps", "1", "--size", "512"])
    assert p.returncode != 0
    assert "needs a GPU" in (p.stderr + p.stdout)
