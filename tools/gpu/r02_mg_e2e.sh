#!/bin/bash
N=$1
mkdir -p gpurun_out
for v in 22 21; do
KZGB200_HOST_PIECE_MIN_LOG=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --prove-log-n 0 --no-sweep > gpurun_out/r02_mg${N}_e2e_$v.json 2> gpurun_out/r02_mg${N}_e2e_$v.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_mg${N}_e2e_$v.json').read().strip().splitlines()[-1])
    print("piece_min_log=$v", {k:d[k] for k in ['value','ms_per_step','n_gpus']}, d['e2e'])
except Exception as e:
    print("bench failed", e); print(open('gpurun_out/r02_mg${N}_e2e_$v.err').read()[-1500:])
PY
done
