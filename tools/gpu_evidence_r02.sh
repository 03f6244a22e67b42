#!/bin/bash
# Round-2 evidence session (1 GPU): tests, smoke, both bench arms (the default command the driver runs), the other two
# workloads, the launch list of the bench and ncu --set full captures of the dominant kernels.  Everything lands in
# gpurun_out/ev_* ; summaries are copied into profiles/r02_* afterwards (tools/collect_evidence_r02.sh).
mkdir -p gpurun_out; rm -f gpurun_out/ev_*
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/ev_smi.csv 2>&1
nproc >> gpurun_out/ev_smi.csv; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/ev_smi.csv
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/ev_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/ev_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ev_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/ev_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/ev_bench_ref.json 2> gpurun_out/ev_bench_ref.err
timeout 900 python bench.py > gpurun_out/ev_bench_n1.json 2> gpurun_out/ev_bench_n1.err; echo "bench rc=$?" >> gpurun_out/ev_bench_n1.err
timeout 600 python bench.py --workload dgemm --steps 3 --warmup 3 --e2e-steps 2 > gpurun_out/ev_bench_dgemm_n1.json 2> gpurun_out/ev_bench_dgemm_n1.err
timeout 600 python bench.py --workload conv --steps 30 --warmup 5 > gpurun_out/ev_bench_conv_n1.json 2> gpurun_out/ev_bench_conv_n1.err
timeout 600 python bench.py --workload conv --steps 30 --warmup 5 --no-graph --no-e2e > gpurun_out/ev_bench_conv_n1_eager.json 2> gpurun_out/ev_bench_conv_n1_eager.err
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --no-verify > gpurun_out/ev_launches_bench.out 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ev_launches_conv.csv python tools/prof_conv.py all 2 > gpurun_out/ev_launches_conv.out 2>&1
prof() {  # name regex skip count cmd...
  local name=$1 regex=$2 skip=$3 count=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s $skip -c $count -f -o gpurun_out/ev_prof_$name "$@" > gpurun_out/ev_ncu_$name.log 2>&1
  ncu -i gpurun_out/ev_prof_$name.ncu-rep --page raw --csv > gpurun_out/ev_prof_$name.raw.csv 2>/dev/null
  ncu -i gpurun_out/ev_prof_$name.ncu-rep --page source --csv > gpurun_out/ev_prof_$name.source.csv 2>/dev/null
  python tools/ncu_top.py gpurun_out/ev_prof_$name.source.csv 30 > gpurun_out/ev_prof_$name.top.txt 2>&1
  rm -f gpurun_out/ev_prof_$name.source.csv gpurun_out/ev_prof_$name.ncu-rep
}
prof tc_bench "gemm_tf32x3" 0 1 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs --no-verify
prof dmma_tma "contract_dmma_tma" 1 1 python tools/profile_kernels.py simt 2
prof i64 "contract_simt_kernel<long" 1 1 python tools/profile_kernels.py simt 2
prof i32 "contract_simt_kernel<int," 1 1 python tools/profile_kernels.py simt 2
prof skinny "skinny_" 2 2 python tools/profile_kernels.py skinny 2
prof conv_c1 "conv_c1" 2 2 python tools/prof_conv.py cv1 3
prof conv_tc_fwd "conv_tc_kernel" 1 1 python tools/prof_conv.py cv2 3
prof conv_tc_dgrad "conv_dgrad_tc_kernel" 1 1 python tools/prof_conv.py cv2 3
prof conv_tc_wgrad "conv_wgrad_tc_kernel" 1 1 python tools/prof_conv.py cv2 3
du -sh gpurun_out; tail -3 gpurun_out/ev_pytest_gpu.txt; tail -2 gpurun_out/ev_smoke.log; cut -c1-300 gpurun_out/ev_bench_n1.json; tail -2 gpurun_out/ev_bench_n1.err
