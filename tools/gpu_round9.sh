#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_tc_kernel" -s 1 -c 1 -f -o gpurun_out/prof_convtc1 python tools/profile_kernels.py convtc 2 > gpurun_out/ncu_convtc1.log 2>&1
ncu -i gpurun_out/prof_convtc1.ncu-rep --page source --csv > gpurun_out/prof_convtc1.source.csv 2>/dev/null
python - <<'PY'
import csv
csv.field_size_limit(1<<30)
rows=list(csv.reader(open('gpurun_out/prof_convtc1.source.csv',errors='replace')))
hdr=None
out=open('gpurun_out/convtc1_sass.txt','w')
for r in rows:
    if any('Sampling' in c for c in r) and 'Source' in r:
        hdr=r; si=r.index('Source'); sa=[i for i,c in enumerate(r) if c.startswith('Warp Stall Sampling (All')][0]; ex=[i for i,c in enumerate(r) if c.startswith('Instructions Executed')][0]
        stall_cols=[(i,c) for i,c in enumerate(r) if c.startswith('stall_')]
        continue
    if hdr and len(r)==len(hdr):
        st=sorted(((float(r[i] or 0),c) for i,c in stall_cols), reverse=True)[:2]
        out.write(f"{r[sa]:>6} {r[ex]:>8} {r[si][:90]:90} {st}\n")
out.close()
PY
rm -f gpurun_out/prof_convtc1.source.csv gpurun_out/prof_convtc1.ncu-rep
grep -n "STG\|ST.E" gpurun_out/convtc1_sass.txt | head -30
