#!/bin/bash
# conv workload on $1 GPUs: CUDA-graph replay vs eager launches
N=$1; mkdir -p gpurun_out
for mode in graph eager; do
  extra=""; [ $mode == eager ] && extra="--no-graph"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --workload conv --steps 30 --warmup 5 --e2e-steps 5 $extra > gpurun_out/r2conv_n${N}_$mode.json 2> gpurun_out/r2conv_n${N}_$mode.err
  echo "$mode rc=$?"; python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2conv_n${N}_$mode.json'))
    print('$mode', round(d['ms_per_step'], 4), 'ms', d['config'].get('launch'), d['parity']['ok'], d['parity']['value'], 'launches', d['gpu_launches'])
except Exception as e:
    print('ERR', e)
PY
  tail -3 gpurun_out/r2conv_n${N}_$mode.err | cut -c1-300
done
