#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/dbg_r2g.py dmma > gpurun_out/r2g_dmma.txt 2>&1
cat gpurun_out/r2g_dmma.txt
timeout 300 compute-sanitizer --tool memcheck python tools/dbg_r2g.py conv > gpurun_out/r2g_conv.txt 2>&1
grep -v "^$" gpurun_out/r2g_conv.txt | head -60
