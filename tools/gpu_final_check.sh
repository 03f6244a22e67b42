#!/bin/bash
# last regression of the round: the two commands the driver runs on the GPU box, plus a short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --configs-only C4,skinny > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/final_pytest_gpu.txt; tail -2 gpurun_out/final_smoke.log; cut -c1-200 gpurun_out/final_bench.json
