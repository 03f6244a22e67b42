#!/bin/bash
# quick multi-GPU check of the fused path: verification at 16384 + one timed run at 32768 (no e2e / cpu legs)
N=$1
mkdir -p gpurun_out
run() { local name=$1 port=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
run mg_verify_$N 29511 --steps 2 --warmup 3 --size 16384 --verify --no-e2e
run bench_n${N}_fused 29512 --steps 6 --warmup 3 $2
for f in gpurun_out/mg_verify_$N.json gpurun_out/bench_n${N}_fused.json; do python -c "
import json
d=json.load(open('$f')); print('$f', round(d['value']/1e3,1), 'TF/s', round(d['ms_per_step'],2), 'ms kern', round(d['roofline']['kernel_ms'],2), d['config'].get('comm'), d.get('verify'), d.get('e2e', {}).get('value'), d['clocks'])
"; done; tail -3 gpurun_out/bench_n${N}_fused.err | cut -c1-300
