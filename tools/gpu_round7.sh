#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "tc" > gpurun_out/pytest_conv_tc.log 2>&1; tail -8 gpurun_out/pytest_conv_tc.log
for g in 1 2 4; do echo "TC groups=$g"; AM_CONVTC_GROUPS=$g AM_BRINGUP_CONV_PATH=3 timeout 120 python tools/gpu_bringup.py conv_speed 2>&1 | cut -c1-900; done
