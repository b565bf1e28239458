#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_large_parity.py -x -q -m gpu -k "ntt or extend or tile" 2>&1 | tail -3
timeout 300 python tools/ops_bench.py 20 22 24 2>&1 | cut -c1-130 | tee gpurun_out/r02_c11_ops.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_c11_bench.json 2> gpurun_out/r02_c11_bench.err; tail -2 gpurun_out/r02_c11_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_c11_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','e2e','gpu_launches']})
r=d['roofline']; print({k:r[k] for k in ['achieved','peak','frac','executed_macs_source','traffic','algorithmic_frac']})
n=d['roofline_ntt']; print({k:n[k] for k in ['achieved','frac','modmul_per_element','ms','traffic','executed_macs_source']})
print(d['prove']['median_ms'], d['prove']['also']['median_ms'])
PY
