#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_round2.py tests/test_gpu_nn_ops.py -m gpu -x -q > gpurun_out/r2b_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.txt
tail -15 gpurun_out/r2b_pytest.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-verify --configs-only C4,skinny > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
grep "cv\|skinny" gpurun_out/r2b_bench.err
