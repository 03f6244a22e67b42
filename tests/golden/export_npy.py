"""Golden-vector exchange in NPY v1 (SURVEY §8c last row): writes every known-answer vector of known_answers.py, plus
a few seeded mid-size cases computed by the CPU oracle, as little-endian C-order .npy files that the reference's own
reader (`read_npy`, src/arraymancer/io/io_npy.nim — header `{'descr': '<f8', 'fortran_order': False, 'shape': (..), }`)
can load, so a Nim build of the reference can consume the same inputs / expected outputs:

    let a = read_npy[int]("tests/golden/npy/int_8x8_8x8.a.npy"); let b = ...; doAssert a * b == read_npy[int](".ab.npy")

    python tests/golden/export_npy.py        # regenerates tests/golden/npy/ and manifest.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests.golden import known_answers as KA  # noqa: E402

OUT = os.path.join(HERE, "npy")
NP = {"f32": "<f4", "f64": "<f8", "i32": "<i4", "i64": "<i8"}


def _save(manifest, case, name, arr, dtype, src):
    a = np.ascontiguousarray(np.asarray(arr, dtype=np.dtype(NP[dtype])))
    fn = f"{case}.{name}.npy"
    with open(os.path.join(OUT, fn), "wb") as f:
        np.lib.format.write_array(f, a, version=(1, 0))
    manifest.setdefault(case, {"src": src, "files": {}})["files"][name] = {"file": fn, "descr": NP[dtype], "shape": list(a.shape)}


def main():
    os.makedirs(OUT, exist_ok=True)
    m = {}
    for c in KA.GEMM:
        for nm in ("a", "b", "ab"):
            _save(m, c["name"], nm, c[nm], c["dtype"], c["src"])
    t = KA.TRANSPOSE
    for nm in ("a", "b", "at", "bt", "expected"):
        _save(m, "transpose_f64", nm, t[nm], "f64", t["src"])
    t = KA.COLMAJOR_SLICE
    for nm in ("a", "eigvecs", "expected"):
        _save(m, "colmajor_reversed_slice_f64", nm, t[nm], "f64", t["src"])
    for key, case in (("CONV_SIMPLE", "conv_simple"), ("CONV_STRIDED", "conv_strided")):
        t = getattr(KA, key)
        for dt in ("i64", "f32"):
            for nm in ("input", "kernel", "bias", "target"):
                arr = np.asarray(t[nm]).reshape(-1, 1, 1) if nm == "bias" else t[nm]
                _save(m, f"{case}_{dt}", nm, arr, dt, t["src"] + f" padding={t['padding']} stride={t['stride']}")
    # seeded mid-size cases, expected values from the oracle (restated laser gemm_strided / im2col conv)
    from oracle import laser_oracle as orc
    from tests.conftest import splitmix64
    orc.build()
    a = (splitmix64(42, 64 * 48) % np.uint64(100)).astype(np.int64).reshape(64, 48)
    b = (splitmix64(43, 48 * 32) % np.uint64(100)).astype(np.int64).reshape(48, 32)
    for nm, arr in (("a", a), ("b", b), ("ab", orc.matmul(a, b))):
        _save(m, "seeded_i64_64x48x32", nm, arr, "i64", "splitmix64 seeds 42/43 mod 100 (benchmarks/integer_matmul.nim inputs); oracle")
    a = splitmix64(7, 33 * 21).view(np.int64).reshape(33, 21); b = splitmix64(8, 21 * 17).view(np.int64).reshape(21, 17)
    for nm, arr in (("a", a), ("b", b), ("ab", orc.matmul(a, b))):
        _save(m, "seeded_i64_fullrange_wrap", nm, arr, "i64", "splitmix64 seeds 7/8 full range: products wrap mod 2^64; oracle")
    rng = np.random.default_rng(2024)
    x = rng.random((2, 1, 28, 28)).astype(np.float32); w = (rng.standard_normal((20, 1, 5, 5)) * 0.28).astype(np.float32)
    bias = rng.random((20, 1, 1)).astype(np.float32)
    y = orc.conv2d(x, w, bias)
    go = np.ones_like(y)
    gi, gw, gb = orc.conv2d_backward(x, w, go)
    for nm, arr in (("input", x), ("kernel", w), ("bias", bias), ("output", y), ("grad_input", gi), ("grad_kernel", gw), ("grad_bias", gb)):
        _save(m, "seeded_lenet_cv1_f32", nm, arr, "f32", "examples/ex02 cv1 shape, numpy default_rng(2024), grad_output = ones; oracle")
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(m, f, indent=1, sort_keys=True)
    print(f"wrote {sum(len(v['files']) for v in m.values())} files for {len(m)} cases under {OUT}")


if __name__ == "__main__":
    main()
