#!/bin/bash
# Round evidence session (1 GPU): tests, bench both arms, launch list of the bench, ncu captures of the dominant kernels.
# Everything lands in gpurun_out/ (kept < 64 MiB); the summaries are copied into profiles/ afterwards.
# $1 = "all" also re-captures the SIMT / DMMA kernels (unchanged since the first evidence session of the round).
mkdir -p gpurun_out; rm -f gpurun_out/prof_* gpurun_out/*.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.csv 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
python tools/gpu_bringup.py peaks peaks_small_n simt_speed f64_speed tc2_speed_f2 conv_speed nn_speed > gpurun_out/bringup_final.log 2>&1
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_bench.out 2>&1
prof() {  # name regex skip count cmd...
  local name=$1 regex=$2 skip=$3 count=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s $skip -c $count -f -o gpurun_out/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/prof_$name.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv > gpurun_out/prof_$name.source.csv 2>/dev/null
  python tools/ncu_top.py gpurun_out/prof_$name.source.csv 30 > gpurun_out/prof_$name.top.txt 2>&1
  rm -f gpurun_out/prof_$name.source.csv
  sz=$(stat -c %s gpurun_out/prof_$name.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 8000000 ]; then rm -f gpurun_out/prof_$name.ncu-rep; fi
}
prof tc_bench "gemm_tf32x3" 0 1 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu
prof conv_tc_fwd "conv_tc_kernel" 0 1 python tools/profile_kernels.py conv 1
prof conv_tc_dgrad "conv_dgrad_tc_kernel" 0 1 python tools/profile_kernels.py conv 1
prof conv_tc_wgrad "conv_wgrad_tc_kernel" 1 1 python tools/profile_kernels.py conv 1
if [ "$1" == "all" ]; then
prof i64wide "contract_simt_kernel<long" 0 1 python tools/profile_kernels.py i64wide 1
prof i32 "contract_simt_kernel<int," 0 1 python tools/profile_kernels.py simt 1
prof dmma "contract_dmma" 0 1 python tools/profile_kernels.py simt 1
prof conv_direct "conv_direct" 0 2 python tools/profile_kernels.py conv 1
fi
du -sh gpurun_out; tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench_n1.json
