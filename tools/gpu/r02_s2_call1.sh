#!/bin/bash
# session 2, call 1: the linked host-scalar pieces (one shared bucket reduction): tests + e2e A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm_multi.py tests/test_gpu_mgpu.py -x -q -m gpu > gpurun_out/r02_s2c1_tests.log 2>&1; tail -4 gpurun_out/r02_s2c1_tests.log
for n in 21 22 24; do for l in 1 0; do
echo "== 2^$n host_link=$l"; KZGB200_HOST_LINK=$l timeout 200 python tools/mgpu_bench.py $n 0 2>&1 | grep e2e
done; done 2>&1 | tee gpurun_out/r02_s2c1_link.log
