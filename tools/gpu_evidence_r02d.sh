#!/bin/bash
# last evidence pass: tests, smoke, conv timings, bench with the conv / skinny records, conv launch list, wgrad ncu
mkdir -p gpurun_out; rm -f gpurun_out/ev4_*
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/ev4_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/ev4_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ev4_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/ev4_smoke.log
timeout 200 python tools/time_conv.py all > gpurun_out/ev4_time_conv.txt 2>&1
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --configs-only C4,skinny > gpurun_out/ev4_bench_c4.json 2> gpurun_out/ev4_bench_c4.err; echo "bench rc=$?"
timeout 600 python bench.py --workload conv --steps 30 --warmup 5 > gpurun_out/ev4_bench_conv_n1.json 2> gpurun_out/ev4_bench_conv_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ev4_launches_conv.csv python tools/prof_conv.py all 2 > gpurun_out/ev4_launches_conv.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_wgrad_tc_kernel" -s 1 -c 1 -f -o gpurun_out/ev4_prof_wgrad python tools/prof_conv.py cv2 3 > gpurun_out/ev4_ncu_wgrad.log 2>&1
ncu -i gpurun_out/ev4_prof_wgrad.ncu-rep --page raw --csv > gpurun_out/ev4_prof_wgrad.raw.csv 2>/dev/null
ncu -i gpurun_out/ev4_prof_wgrad.ncu-rep --page source --csv > gpurun_out/ev4_prof_wgrad.source.csv 2>/dev/null
python tools/ncu_top.py gpurun_out/ev4_prof_wgrad.source.csv 30 > gpurun_out/ev4_prof_wgrad.top.txt 2>&1
rm -f gpurun_out/ev4_prof_wgrad.source.csv gpurun_out/ev4_prof_wgrad.ncu-rep
tail -3 gpurun_out/ev4_pytest_gpu.txt; tail -2 gpurun_out/ev4_smoke.log; cat gpurun_out/ev4_time_conv.txt
