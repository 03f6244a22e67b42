#!/bin/bash
# multi-GPU session: $1 = number of GPUs.  Fused GEMM + peer stores (default) next to the NCCL all-gather variant.
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
run() {  # name port args...
  local name=$1 port=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"
}
run mg_verify_$N 29511 --steps 2 --warmup 3 --size 16384 --verify --no-e2e
run bench_n$N 29512 --steps 6 --warmup 3
run bench_n${N}_nccl_c1 29513 --steps 4 --warmup 3 --comm nccl --chunks 1 --no-e2e
if [ "$2" == "more" ]; then run bench_n${N}_nccl_c4 29514 --steps 4 --warmup 3 --comm nccl --chunks 4 --no-e2e; fi
for f in gpurun_out/mg_verify_$N.json gpurun_out/bench_n$N*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(round(d['value']/1e3,1), 'TF/s', round(d['ms_per_step'],2), 'ms kern', round(d['roofline']['kernel_ms'],2), 'e2e', d.get('e2e',{}).get('value'), d['config'].get('comm'), d.get('verify'), d['clocks'])
except Exception as e: print('ERR',e)
"; done; tail -4 gpurun_out/bench_n$N.err | cut -c1-300
