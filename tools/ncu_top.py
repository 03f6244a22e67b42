"""Summarise an `ncu --page source --csv` export: per-kernel top instructions by warp-stall samples and the
sample share per opcode.  usage: python tools/ncu_top.py file.csv [topN]"""
import csv
import sys
from collections import defaultdict

csv.field_size_limit(1 << 30)
path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path, errors="replace")))
# the export is a sequence of tables; find header rows (contain "Source" and "Sampling")
i = 0
while i < len(rows):
    r = rows[i]
    if any("Sampling" in c for c in r) and any(c.strip() in ("Source", "SASS") or "Source" in c for c in r):
        hdr = r
        col = {c: j for j, c in enumerate(hdr)}
        src_c = next((j for j, c in enumerate(hdr) if c.strip() in ("Source", "SASS")), None)
        samp_c = next((j for j, c in enumerate(hdr) if c.startswith("# Samples") or c.startswith("Warp Stall Sampling (All")), None)
        ni_c = next((j for j, c in enumerate(hdr) if c.startswith("Warp Stall Sampling (Not-issued") or c.startswith("# Samples (Not")), None)
        exec_c = next((j for j, c in enumerate(hdr) if c.startswith("Instructions Executed") or c.startswith("# Instructions Executed")), None)
        body = []
        i += 1
        while i < len(rows) and not (any("Sampling" in c for c in rows[i]) and len(rows[i]) == len(hdr) and rows[i][0] == hdr[0]):
            if len(rows[i]) == len(hdr):
                body.append(rows[i])
            i += 1
        def num(x):
            try:
                return float(x.replace(",", ""))
            except Exception:
                return 0.0
        tot = sum(num(b[samp_c]) for b in body) or 1.0
        print(f"=== table with {len(body)} instructions, total samples {tot:.0f}; columns: {hdr[:12]}")
        byop = defaultdict(float)
        for b in body:
            op = b[src_c].strip().split()[0] if b[src_c].strip() else "?"
            if op.startswith("@"):
                parts = b[src_c].strip().split()
                op = parts[1] if len(parts) > 1 else op
            byop[op] += num(b[samp_c])
        print("samples by opcode:", ", ".join(f"{k}:{v / tot:.3f}" for k, v in sorted(byop.items(), key=lambda kv: -kv[1])[:14]))
        for b in sorted(body, key=lambda b: -num(b[samp_c]))[:topn]:
            print(f"  {num(b[samp_c]) / tot:6.3f}  exec={b[exec_c] if exec_c is not None else '':>10}  {b[src_c].strip()[:110]}")
        continue
    i += 1
