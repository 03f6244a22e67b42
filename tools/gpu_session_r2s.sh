#!/bin/bash
python - <<'PY' 2>&1 | tee gpurun_out/r2s.txt
import sys; sys.path.insert(0, '.')
from arraymancer_b200 import _capi
for w, name in ((31, "issue pattern of the conv kernels (3xTF32, commits, waits, chains)"), (32, "+ 16 KB per k block streamed from L2"), (33, "pattern + operand-stage writers (tcgen05.st)"), (34, "pattern + accumulator readers (tcgen05.ld)"), (35, "all")):
    v = _capi.microbench(w)
    print(w, name, round(v, 1), "TFLOP/s  ->", round(2 * 128 * 64 * 8 / (v * 1e12 / 148) * 1.92e9, 1), "cycles per MMA at 1.92 GHz")
PY
