#!/bin/bash
mkdir -p gpurun_out
prof() {  # name regex count script-arg
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -c $3 -f -o gpurun_out/prof_$1 python tools/profile_kernels.py $4 1 > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$1.ncu-rep --page source --csv > gpurun_out/prof_$1.source.csv 2>/dev/null
  python tools/ncu_top.py gpurun_out/prof_$1.source.csv 30 > gpurun_out/prof_$1.top.txt 2>&1
  gzip -f gpurun_out/prof_$1.source.csv
  sz=$(stat -c %s gpurun_out/prof_$1.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 9000000 ]; then rm -f gpurun_out/prof_$1.ncu-rep; fi
}
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python tools/gpu_bringup.py conv_speed 2>&1 | cut -c1-700
prof convd "conv_direct" 4 conv
prof i32 "contract_simt_kernel<int," 1 simt
prof i64 "contract_simt_kernel<long" 1 simt
head -c 2500 gpurun_out/prof_convd.top.txt; head -c 2500 gpurun_out/prof_i32.top.txt
du -sh gpurun_out
