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
