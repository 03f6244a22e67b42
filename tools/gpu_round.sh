#!/bin/bash
# One GPU-box session: tests, bench (both arms), launch list and ncu captures.  Output -> gpurun_out/ (<= 64 MiB).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.csv 2>&1
nproc > gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt; free -g >> gpurun_out/host.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
if [ "$1" != "nobench" ]; then
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
fi
python tools/gpu_bringup.py f64_speed > gpurun_out/bringup_f64.log 2>&1
prof() {  # name regex count script-arg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -c $3 -f -o gpurun_out/prof_$1 python tools/profile_kernels.py $4 1 > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$1.ncu-rep --page source --csv > gpurun_out/prof_$1.source.csv 2>/dev/null
  ls -la gpurun_out/prof_$1.* >> gpurun_out/sizes.txt
  sz=$(stat -c %s gpurun_out/prof_$1.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 12000000 ]; then rm -f gpurun_out/prof_$1.ncu-rep; fi
  sz=$(stat -c %s gpurun_out/prof_$1.source.csv 2>/dev/null || echo 0)
  if [ "$sz" -gt 12000000 ]; then gzip -f gpurun_out/prof_$1.source.csv; fi
}
prof simt "contract_simt|contract_dmma" 5 simt
prof conv contract_simt 6 conv
prof tc gemm_tf32x3 1 tc
du -sh gpurun_out >> gpurun_out/sizes.txt
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_n1.json 2>/dev/null | cut -c1-600; cat gpurun_out/bringup_f64.log | cut -c1-1500
