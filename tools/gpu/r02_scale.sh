#!/bin/bash
# usage: r02_scale.sh N  -- the driver's bench line at N GPUs (reference arm first when N = 1)
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_scaling_${N}gpu.json 2> gpurun_out/r02_scaling_${N}gpu.err
  timeout 300 python tools/mgpu_bench.py 24 $(python -c "print(','.join(str(i) for i in range($N)))") > gpurun_out/r02_mgpu_${N}gpu.log 2>&1
  tail -2 gpurun_out/r02_mgpu_${N}gpu.log
fi
python - <<PY
import json,glob
f='gpurun_out/r02_bench_1gpu.json' if "$N"=="1" else 'gpurun_out/r02_scaling_${N}gpu.json'
d=json.loads(open(f).read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','n_gpus','gpu_launches']}, d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])
r=d['roofline']; print({k:r[k] for k in ['achieved','peak','frac','executed_macs_source','traffic','algorithmic_frac','kernel_ms_per_launch']})
if d.get('prove'): print(d['prove']['median_ms'], d['prove'].get('also',{}).get('median_ms'), d['prove'].get('byte_identical_to_cpu_oracle'))
if d.get('roofline_ntt'): print({k:d['roofline_ntt'][k] for k in ['ms','frac','modmul_per_element','traffic']})
PY
