#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nn_ops.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2e_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.txt
tail -12 gpurun_out/r2e_pytest.txt
ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.sum --clock-control none --csv --log-file gpurun_out/r2e_launches.csv python tools/prof_conv.py cv1 3 > gpurun_out/r2e_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_c1 -s 2 -c 2 -o gpurun_out/r2e_c1 python tools/prof_conv.py cv1 3 > gpurun_out/r2e_p.log 2>&1
echo done
