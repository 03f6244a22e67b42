#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/gpu_bringup.py peaks simt_speed f64_speed conv_speed > gpurun_out/bringup3.log 2>&1
timeout 600 python bench.py --steps 4 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
tail -12 gpurun_out/pytest_gpu.log; cut -c1-1200 gpurun_out/bringup3.log; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
