#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/time_conv.py cv2 2>&1 | tee gpurun_out/r2t_time.txt
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nn_ops.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2t_pytest.txt
timeout 150 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2t_launches.csv python tools/prof_conv.py cv2 2 > gpurun_out/r2t_l.log 2>&1
python tools/ncu_launches.py gpurun_out/r2t_launches.csv tc | tail -6
