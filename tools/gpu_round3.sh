#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
python tools/gpu_bringup.py conv_speed simt_speed 2>&1 | cut -c1-1800
