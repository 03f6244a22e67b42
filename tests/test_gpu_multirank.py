"""GPU, world size 2 (one process per GPU under torchrun): the sharded paths of arraymancer_b200.distributed on real
NCCL / NVLink — fused GEMM + all-gather epilogue over symmetric memory, NCCL all-gather for float64 / int64, the
host-buffer sharded GEMM, batch-sharded conv.  Skipped on a one-GPU box (the driver's GPU test box); run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world2_sharded_paths():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", os.path.join(ROOT, "tests", "mg_worker.py")], capture_output=True, text=True,
                       timeout=900, env=env)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-4000:])
    assert "MG_OK 0" in p.stdout and "MG_OK 1" in p.stdout
