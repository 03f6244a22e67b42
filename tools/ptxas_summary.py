"""Compact per-kernel register / spill / smem summary from the -Xptxas -v logs of the last build
(arraymancer_b200/csrc/build/*.ptxas.log).  Usage: python tools/ptxas_summary.py [file-substring]"""
import glob, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "arraymancer_b200", "csrc", "build")
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for f in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    if pat not in os.path.basename(f):
        continue
    txt = open(f).read()
    for m in re.finditer(r"Compiling entry function '([^']+)'.*?\n(.*?)(?=ptxas info\s*: Compiling|\Z)", txt, re.S):
        name, body = m.group(1), m.group(2)
        regs = re.search(r"Used (\d+) registers", body)
        spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", body)
        smem = re.search(r"(\d+) bytes smem", body)
        try:
            dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip()
        except Exception:
            dem = name
        dem = re.sub(r"am::", "", dem)
        dem = dem[:150]
        print(f"{os.path.basename(f)[:-10]:16s} regs={regs.group(1) if regs else '?':>3} spill={spill.group(1) if spill else '?':>4}/{spill.group(2) if spill else '?':<4} smem={smem.group(1) if smem else '0':>6}  {dem}")
