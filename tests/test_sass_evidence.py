"""CPU: the built library really contains the Blackwell instructions the design claims (no GPU needed: SASS of the
sm_100a cubin via cuobjdump).  tcgen05.mma = UTCHMMA, TMA = UTMALDG / UBLKCP, tcgen05.ld/st = LDTM / STTM,
mma.sync f64 = DMMA, cp.async = LDGSTS."""
import collections
import os
import re
import shutil
import subprocess

import pytest

from arraymancer_b200 import _capi

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
INSN = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z][A-Za-z0-9_.]+)")


@pytest.fixture(scope="module")
def sass():
    if not os.path.exists(CUOBJDUMP) or not os.path.exists(_capi.LIB_PATH):
        pytest.skip("cuobjdump or the built library is missing")
    out = subprocess.run([CUOBJDUMP, "-sass", _capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in out.splitlines():
        if line.lstrip().startswith("Function :"):
            cur = line.split(":", 1)[1].strip()
            kernels[cur] = collections.Counter()
        else:
            m = INSN.match(line)
            if m and cur:
                kernels[cur][m.group(1)] += 1
    return kernels


def _count(c, prefix):
    return sum(v for k, v in c.items() if k == prefix or k.startswith(prefix + ".") or k.startswith(prefix))


def _find(kernels, needle):
    hits = [(n, c) for n, c in kernels.items() if needle in n]
    assert hits, f"no kernel matching {needle}"
    return hits


def test_gemm_f32_uses_tcgen05_tma_tmem(sass):
    for name, c in _find(sass, "gemm_tf32x3_kernel"):
        assert _count(c, "UTCHMMA") >= 12 and _count(c, "UTMALDG") >= 4 and _count(c, "LDTM") >= 1 and _count(c, "UTCBAR") >= 1, name


def test_conv_kernels_use_tcgen05_with_operands_in_tensor_memory(sass):
    for needle in ("conv_tc_kernel", "conv_dgrad_tc_kernel", "conv_wgrad_tc_kernel"):
        for name, c in _find(sass, needle):
            assert _count(c, "UTCHMMA") >= 3, name            # tcgen05.mma
            assert _count(c, "STTM") >= 2, name               # tcgen05.st: the A operand is written to TMEM
            assert _count(c, "LDTM") >= 1, name               # accumulators read back with tcgen05.ld
    for name, c in _find(sass, "conv_tc_kernel"):
        assert _count(c, "UBLKCP") >= 1 and _count(c, "UTMALDG") >= 2, name     # bulk copy of raw images + TMA weight tiles


def test_f64_uses_dmma_and_simt_kernels_use_cp_async(sass):
    for name, c in _find(sass, "contract_dmma_kernel"):
        assert _count(c, "DMMA") >= 32, name
    ints = [c for n, c in sass.items() if "contract_simt_kernel" in n and ("Ii" in n or "Il" in n)]
    assert ints and all(_count(c, "IMAD") > 100 for c in ints)
    assert any(_count(c, "LDGSTS") > 0 for n, c in sass.items() if "contract_simt_kernel" in n)
    for name, c in _find(sass, "conv_direct_f32_kernel"):
        assert _count(c, "LDGSTS") >= 2 and _count(c, "FFMA") >= 32, name


def test_no_heavy_spills(sass):
    for name, c in sass.items():
        assert _count(c, "STL") <= 16, f"{name}: {_count(c, 'STL')} local-memory stores"
