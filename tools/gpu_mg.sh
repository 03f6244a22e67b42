#!/bin/bash
# multi-GPU session: $1 = number of GPUs
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1

timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 --size 16384 --verify --no-e2e > gpurun_out/mg_verify_$N.json 2> gpurun_out/mg_verify_$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
for c in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$c bench.py --gpus $N --steps 4 --warmup 3 --chunks $c --no-e2e > gpurun_out/bench_n${N}_c$c.json 2> gpurun_out/bench_n${N}_c$c.err
done
cat gpurun_out/mg_verify_$N.json | cut -c1-300; tail -2 gpurun_out/mg_verify_$N.err; for f in gpurun_out/bench_n$N*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d.get('e2e',{}).get('value'), d['clocks'])
except Exception as e: print('ERR',e)
"; done; tail -3 gpurun_out/bench_n$N.err
