#!/bin/bash
# session 2, final pass C: bench.py (this repo's arm) at 1 GPU on the final sources
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_s2_bench_1gpu.json 2> gpurun_out/r02_s2_bench_1gpu.err; tail -3 gpurun_out/r02_s2_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_s2_bench_1gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','e2e','gpu_launches']})
r=d['roofline']; print({k:r[k] for k in ['achieved','peak','frac','executed_macs_source','traffic','algorithmic_frac','kernel_ms_per_launch']})
n=d['roofline_ntt']; print({k:n[k] for k in ['achieved','frac','modmul_per_element','ms','traffic','executed_macs_source']})
print(d['prove']['median_ms'], d['prove']['also']['median_ms'], d['prove'].get('byte_identical_to_cpu_oracle'))
PY
