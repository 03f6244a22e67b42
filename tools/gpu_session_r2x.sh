#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/time_conv.py cv2 2>&1 | tee gpurun_out/r2x_time.txt
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nn_ops.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2x_pytest.txt
