#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/pytest_conv.log 2>&1; tail -8 gpurun_out/pytest_conv.log
python tools/gpu_bringup.py conv_speed 2>&1 | cut -c1-900
for g in 1 2 3 4 5 6 -2 -4 -8; do AM_TC_GROUP=$g python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('group $g', round(d['value']), round(d['roofline']['kernel_ms'],1), d['clocks']['sm_mhz'])"; done
