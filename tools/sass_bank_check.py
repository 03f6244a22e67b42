"""Static register-bank check of the FFMAs of one kernel: on sm_100 a 3-source FFMA issues in one cycle only when the
sources it actually reads from the register file (those not served by the operand reuse cache) sit in different banks
(even / odd register index).  Prints the share of FFMAs with a same-bank pair.  usage: sass_bank_check.py <so> <kernel-substr>"""
import re, subprocess, sys
so, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, stats = None, {}
prev_reuse = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); prev_reuse = {}
        continue
    if cur is None or pat not in cur:
        continue
    m = re.search(r"\bFFMA\s+(R\d+|RZ), (-?\|?R\d+\|?|RZ)(\.reuse)?, (-?\|?R\d+\|?|RZ|c\[.*?\]|-?[0-9.e+\-]+|UR\d+)(\.reuse)?, (-?\|?R\d+\|?|RZ)(\.reuse)?", line)
    if not m:
        continue
    srcs = [(m.group(2), m.group(3)), (m.group(4), m.group(5)), (m.group(6), m.group(7))]
    regs = []
    for slot, (r, reuse) in enumerate(srcs):
        rr = re.sub(r"[-|]", "", r)
        if rr.startswith("R") and rr != "RZ":
            n = int(rr[1:])
            cached = prev_reuse.get(slot) == n
            if not cached:
                regs.append(n)
    new_reuse = {}
    for slot, (r, reuse) in enumerate(srcs):
        rr = re.sub(r"[-|]", "", r)
        if reuse and rr.startswith("R") and rr != "RZ":
            new_reuse[slot] = int(rr[1:])
    prev_reuse = new_reuse
    st = stats.setdefault(cur, [0, 0, 0])
    st[0] += 1
    par = [n & 1 for n in regs]
    if len(regs) >= 2 and (par.count(0) >= 2 or par.count(1) >= 2):
        st[1] += 1
    if len(regs) == 3:
        st[2] += 1
for k, (n, conf, three) in stats.items():
    print(f"{n:6d} FFMA  same-parity pair {conf / max(n,1):.2f}  three register reads {three / max(n,1):.2f}  {k[:90]}")
