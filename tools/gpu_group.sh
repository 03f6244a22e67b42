#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_bringup.py peaks 2>&1 | cut -c1-600
for g in 2 4 8 16 32; do
  AM_TC_GROUP=$g timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:gemm_tf32x3" -c 1 --csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' -v g=$g '{print "group",g,$(NF-2),$(NF-1),$NF}'
done
for g in 2 4 8 16 32; do AM_TC_GROUP=$g python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('group $g', round(d['value']), d['roofline']['kernel_ms'], d['clocks'])"; done
