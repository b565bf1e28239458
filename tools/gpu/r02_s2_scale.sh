#!/bin/bash
# usage: r02_s2_scale.sh N -- bench.py at N GPUs (one rank per GPU, no prover / sweep legs) and the one-process C-level form
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --prove-log-n 0 --no-sweep > gpurun_out/r02_s2_scaling_${N}gpu.json 2> gpurun_out/r02_s2_scaling_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_s2_scaling_${N}gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ['value','ms_per_step','n_gpus','gpu_launches']}, d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])
except Exception as e:
    print("bench failed", e); print(open('gpurun_out/r02_s2_scaling_${N}gpu.err').read()[-2000:])
PY
timeout 300 python tools/mgpu_bench.py 24 $(python -c "print(','.join(str(i) for i in range($N)))") > gpurun_out/r02_s2_mgpu_${N}gpu.log 2>&1; tail -2 gpurun_out/r02_s2_mgpu_${N}gpu.log
