"""Tabulate an `ncu --metrics ... --csv --log-file` launch list: one line per launch.  usage: ncu_launches.py file.csv [filter]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
flt = sys.argv[2] if len(sys.argv) > 2 else ""
hdr, agg = None, {}
for r in rows:
    if 'Kernel Name' in r:
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        agg.setdefault((int(d['ID']), d['Kernel Name'][:70]), {})[d['Metric Name']] = d['Metric Value']
for (i, k), m in sorted(agg.items()):
    if flt not in k or any(x in k for x in ('memset', 'FillFunctor', 'distribution')):
        continue
    print(i, k, {a.split('.')[0][-22:]: b for a, b in m.items()})
