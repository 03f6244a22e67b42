#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/pytest_conv.log 2>&1; tail -8 gpurun_out/pytest_conv.log
for kb in 12 16 24 40; do echo "SMEMKB=$kb"; AM_CONV_SMEMKB=$kb python tools/gpu_bringup.py conv_speed 2>&1 | cut -c1-900; done
