#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/convtc_launches.csv python tools/profile_kernels.py convtc 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/convtc_launches.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
for r in rows[1:]: print(r[ik][:60], r[iv], r[iu])
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_tc_kernel" -s 2 -c 1 -f -o gpurun_out/prof_convtc python tools/profile_kernels.py convtc 2 > gpurun_out/ncu_convtc.log 2>&1
ncu -i gpurun_out/prof_convtc.ncu-rep --page raw --csv > gpurun_out/prof_convtc.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_convtc.ncu-rep --page source --csv > gpurun_out/prof_convtc.source.csv 2>/dev/null
python tools/ncu_top.py gpurun_out/prof_convtc.source.csv 40 > gpurun_out/prof_convtc.top.txt 2>&1
rm -f gpurun_out/prof_convtc.source.csv gpurun_out/prof_convtc.ncu-rep
cat gpurun_out/prof_convtc.top.txt | cut -c1-220
