#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/dbg_r2g.py dmma2 2>&1 | tail -8
timeout 300 python tools/dbg_r2g.py convs 2>&1 | tail -12
