#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.txt
tail -12 gpurun_out/r2i_pytest.txt
timeout 300 python tools/dbg_r2g.py dmma 2>&1 | tail -12
timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-verify --configs-only C1,C2,C4,skinny > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
grep "bench\]" gpurun_out/r2i_bench.err | cut -c1-200
ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.sum --clock-control none --csv --log-file gpurun_out/r2i_launches.csv python tools/prof_conv.py all 3 > gpurun_out/r2i_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_c1 -s 2 -c 2 -o gpurun_out/r2i_c1 python tools/prof_conv.py cv1 3 > gpurun_out/r2i_p.log 2>&1
echo done
