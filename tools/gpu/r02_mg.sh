#!/bin/bash
# usage: r02_mg.sh N
N=$1
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_mgpu.py "tests/test_gpu_msm_multi.py::test_two_contexts_on_two_devices" -x -q -m gpu > gpurun_out/r02_mg${N}_tests.log 2>&1; tail -4 gpurun_out/r02_mg${N}_tests.log
DEVS=$(python -c "print(','.join(str(i) for i in range($N)))")
timeout 300 python tools/mgpu_bench.py 24 $DEVS 2>&1 | tee gpurun_out/r02_mg${N}_mgpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --prove-log-n 0 --no-sweep > gpurun_out/r02_mg${N}_bench.json 2> gpurun_out/r02_mg${N}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_mg${N}_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ['value','ms_per_step','e2e','n_gpus','gpu_launches']})
    print(d['roofline']['kernel_ms_per_launch'], d['roofline']['affine_rounds'])
except Exception as e:
    print("bench failed", e); print(open('gpurun_out/r02_mg${N}_bench.err').read()[-2000:])
PY
