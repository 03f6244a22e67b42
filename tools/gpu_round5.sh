#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/pytest_gemm.log 2>&1; tail -5 gpurun_out/pytest_gemm.log
python tools/gpu_bringup.py tc2_speed_f2 tc1_speed_f2 2>&1 | cut -c1-700
for sync in 1 0; do for g in 4 8 9 -8 -4; do AM_TC_SYNC=$sync AM_TC_GROUP=$g python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sync $sync group $g', round(d['value']), round(d['roofline']['kernel_ms'],1), d['clocks']['sm_mhz'])"; done; done
for g in 4 8; do
  AM_TC_GROUP=$g timeout 600 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:gemm_tf32x3" -c 1 --csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' -v g=$g '{print "sync1 group",g,$(NF-2),$(NF-1),$NF}'
done
