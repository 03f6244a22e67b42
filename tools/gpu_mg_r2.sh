#!/bin/bash
# round-2 multi-GPU session: $1 = GPUs.  World-size-N test of the sharded paths, then the three sharded BASELINE workloads
# (sgemm with its host-buffer e2e leg, dgemm, conv), NVLink byte counters around the fused SGEMM run.
N=$1; TAG=r2mg$N
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
if [ "$2" != "notest" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/mg_worker.py > gpurun_out/${TAG}_worker.txt 2>&1; echo "mg_worker rc=$?"
grep -E "MG_OK|Error|assert" gpurun_out/${TAG}_worker.txt | head -12
fi
run() {  # name port args...
  local name=$1 port=$2; shift 2
  timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  echo "$name rc=$?"
}
nvidia-smi nvlink -gt d -i 0 > gpurun_out/${TAG}_nvlink_before.txt 2>&1
run sgemm 29512 --steps 6 --warmup 3 --e2e-steps 4
nvidia-smi nvlink -gt d -i 0 > gpurun_out/${TAG}_nvlink_after.txt 2>&1
run dgemm 29513 --workload dgemm --steps 3 --warmup 3 --e2e-steps 2
run conv 29514 --workload conv --steps 20 --warmup 5 --e2e-steps 5
for w in sgemm dgemm conv; do f=gpurun_out/${TAG}_$w.json; python - <<PY
import json
try:
    d = json.load(open('$f'))
    print('$w', round(d['value'] / 1e3, 2), 'TF/s', round(d['ms_per_step'], 3), 'ms', 'e2e', (d.get('e2e') or {}).get('value'), 'verify', d.get('verify') or d.get('parity'), d['clocks'])
except Exception as e:
    print('$w ERR', e)
PY
tail -3 gpurun_out/${TAG}_$w.err | cut -c1-400
done
