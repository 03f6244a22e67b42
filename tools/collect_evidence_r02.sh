#!/bin/bash
# copies the summaries of the round-2 evidence session (gpurun_out/ev_*) into profiles/ (tracked)
set -e
cd "$(dirname "$0")/.."
mkdir -p profiles/r02_ncu
cp gpurun_out/ev_bench_n1.json profiles/r02_bench_n1.json
cp gpurun_out/ev_bench_ref.json profiles/r02_bench_ref.json
cp gpurun_out/ev_bench_dgemm_n1.json profiles/r02_bench_dgemm_n1.json
cp gpurun_out/ev_bench_conv_n1.json profiles/r02_bench_conv_n1.json
cp gpurun_out/ev_bench_conv_n1_eager.json profiles/r02_bench_conv_n1_eager.json
cp gpurun_out/ev_pytest_gpu.txt profiles/r02_pytest_gpu.txt
cp gpurun_out/ev_smoke.log profiles/r02_smoke.log
cp gpurun_out/ev_smi.csv profiles/r02_box.csv
cp gpurun_out/ev_launches_bench.csv profiles/r02_launches_bench.csv
cp gpurun_out/ev_launches_conv.csv profiles/r02_launches_conv.csv
for f in gpurun_out/ev_prof_*.raw.csv gpurun_out/ev_prof_*.top.txt; do
  [ -s "$f" ] && cp "$f" profiles/r02_ncu/$(basename "$f" | sed 's/^ev_prof_//')
done
python - <<'PY'
import csv, json, os
# DRAM bytes per launch of the bench GEMM from the ncu raw page -> profiles/traffic.json (bench.py's roofline.traffic)
p = "profiles/r02_ncu/tc_bench.raw.csv"
if os.path.exists(p):
    rows = list(csv.reader(open(p)))
    hdr, units = rows[0], rows[1]
    d = dict(zip(hdr, rows[2]))
    def val(k):
        v = float(d[k]); u = units[hdr.index(k)].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
    tot = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    t = json.load(open("profiles/traffic.json")) if os.path.exists("profiles/traffic.json") else {}
    t.setdefault("gemm_tf32x3_kernel", {})["n32768_g1"] = int(tot)
    t["gemm_tf32x3_kernel"]["source_r02"] = "profiles/r02_ncu/tc_bench.raw.csv: dram__bytes_read.sum + dram__bytes_write.sum, one launch"
    json.dump(t, open("profiles/traffic.json", "w"), indent=1)
    print("traffic", tot / 1e9, "GB")
PY
ls profiles/r02_ncu | head -30
