"""Opcode histogram (or listing) of one kernel from the built library.  usage: sass_fn.py <name-substring> [list]"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("AM_B200_LIB") or os.path.join(ROOT, "arraymancer_b200", "libarraymancer_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat, cur, keep = sys.argv[1], None, []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        if pat in cur: keep.append(("F", cur))
        continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m: keep.append(("I", m.group(2).strip()))
if len(sys.argv) > 2:
    for k, v in keep: print(v)
else:
    h = None
    for k, v in keep + [("F", None)]:
        if k == "F":
            if h is not None: print(name, sum(h.values()), dict(h.most_common(14)))
            h = collections.Counter(); name = v
        else:
            op = v.split()[1] if v.startswith("@") else v.split()[0]
            h[op.split(".")[0]] += 1
