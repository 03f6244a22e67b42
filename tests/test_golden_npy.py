"""CPU: the NPY golden-vector exchange (tests/golden/npy): every file parses with a reader that accepts
exactly what the reference's read_npy accepts (io_npy.nim:13-60: v1 header matched by
`{'descr': '$+', 'fortran_order': $+, 'shape': $+, }`), and the contents equal known_answers.py / the oracle."""
import json
import os
import re
import struct

import numpy as np

from tests.golden import known_answers as KA

NPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "npy")


def read_npy_like_reference(path):
    raw = open(path, "rb").read()
    assert raw[:6] == b"\x93NUMPY" and raw[6] == 1 and raw[7] == 0, "NPY v1.0 expected"
    (hlen,) = struct.unpack("<H", raw[8:10])
    header = raw[10:10 + hlen].decode("latin1")
    m = re.match(r"\{'descr': '(.+?)', 'fortran_order': (.+?), 'shape': (.+?), \}", header)
    assert m, header
    descr, fortran, shape = m.group(1), m.group(2), m.group(3)
    assert descr in ("<f4", "<f8", "<i4", "<i8") and fortran == "False"
    dims = tuple(int(x) for x in re.findall(r"\d+", shape))
    return np.frombuffer(raw[10 + hlen:], dtype=np.dtype(descr)).reshape(dims)


def test_manifest_files_parse_and_match_known_answers():
    man = json.load(open(os.path.join(NPY, "manifest.json")))
    n = 0
    for case, ent in man.items():
        for name, f in ent["files"].items():
            arr = read_npy_like_reference(os.path.join(NPY, f["file"]))
            assert list(arr.shape) == f["shape"] and arr.dtype.str == f["descr"]
            n += 1
    assert n >= 60
    for c in KA.GEMM:
        a, b, ab = (read_npy_like_reference(os.path.join(NPY, f"{c['name']}.{k}.npy")) for k in ("a", "b", "ab"))
        assert np.array_equal(a @ b, ab) and np.array_equal(ab, np.asarray(c["ab"]))


def test_seeded_cases_match_the_oracle(oracle):
    a, b, ab = (read_npy_like_reference(os.path.join(NPY, f"seeded_i64_fullrange_wrap.{k}.npy")) for k in ("a", "b", "ab"))
    assert np.array_equal(oracle.matmul(a.copy(), b.copy()), ab)
    g = {k: read_npy_like_reference(os.path.join(NPY, f"seeded_lenet_cv1_f32.{k}.npy"))
         for k in ("input", "kernel", "bias", "output", "grad_input", "grad_kernel", "grad_bias")}
    assert np.array_equal(oracle.conv2d(g["input"].copy(), g["kernel"].copy(), g["bias"].copy()), g["output"])
    gi, gw, gb = oracle.conv2d_backward(g["input"].copy(), g["kernel"].copy(), np.ones_like(g["output"]))
    assert np.array_equal(gi, g["grad_input"]) and np.array_equal(gw, g["grad_kernel"]) and np.array_equal(gb, g["grad_bias"])
