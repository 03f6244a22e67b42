#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2c_launches.csv python tools/prof_conv.py all 3 > gpurun_out/r2c_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_c1 -s 2 -c 2 -o gpurun_out/r2c_c1 python tools/prof_conv.py cv1 3 > gpurun_out/r2c_p.log 2>&1
echo done
