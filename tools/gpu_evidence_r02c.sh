#!/bin/bash
# final evidence pass of round 2: tests, smoke, bench (all configs), conv workload, conv launch list, wgrad + skinny ncu
mkdir -p gpurun_out; rm -f gpurun_out/ev3_*
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/ev3_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/ev3_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ev3_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/ev3_smoke.log
timeout 200 python tools/time_conv.py all > gpurun_out/ev3_time_conv.txt 2>&1
timeout 900 python bench.py > gpurun_out/ev3_bench_n1.json 2> gpurun_out/ev3_bench_n1.err; echo "bench rc=$?" >> gpurun_out/ev3_bench_n1.err
timeout 600 python bench.py --workload conv --steps 30 --warmup 5 > gpurun_out/ev3_bench_conv_n1.json 2> gpurun_out/ev3_bench_conv_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ev3_launches_conv.csv python tools/prof_conv.py all 2 > gpurun_out/ev3_launches_conv.out 2>&1
prof() {  # name regex skip count cmd...
  local name=$1 regex=$2 skip=$3 count=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s $skip -c $count -f -o gpurun_out/ev3_prof_$name "$@" > gpurun_out/ev3_ncu_$name.log 2>&1
  ncu -i gpurun_out/ev3_prof_$name.ncu-rep --page raw --csv > gpurun_out/ev3_prof_$name.raw.csv 2>/dev/null
  ncu -i gpurun_out/ev3_prof_$name.ncu-rep --page source --csv > gpurun_out/ev3_prof_$name.source.csv 2>/dev/null
  python tools/ncu_top.py gpurun_out/ev3_prof_$name.source.csv 30 > gpurun_out/ev3_prof_$name.top.txt 2>&1
  rm -f gpurun_out/ev3_prof_$name.source.csv gpurun_out/ev3_prof_$name.ncu-rep
}
prof skinny "skinny_" 0 6 python tools/profile_kernels.py skinny 1
prof conv_tc_wgrad "conv_wgrad_tc_kernel" 1 1 python tools/prof_conv.py cv2 3
tail -3 gpurun_out/ev3_pytest_gpu.txt; tail -2 gpurun_out/ev3_smoke.log; cat gpurun_out/ev3_time_conv.txt; tail -2 gpurun_out/ev3_bench_n1.err
