#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/time_conv.py cv2 2>&1 | tee gpurun_out/r2u_time.txt
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nn_ops.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2u_pytest.txt
timeout 300 python bench.py --workload conv --steps 30 --warmup 5 --e2e-steps 3 > gpurun_out/r2u_conv_n1.json 2> gpurun_out/r2u_conv_n1.err; echo "conv rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2u_conv_n1.json'))
print(round(d['ms_per_step'], 4), 'ms', d['config'].get('launch'), d['parity']['ok'], d['parity']['value'], 'launches', d['gpu_launches'])
PY
tail -3 gpurun_out/r2u_conv_n1.err | cut -c1-300
