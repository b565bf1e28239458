#!/bin/bash
# session 2, call 8: pieces merged for free (a later piece opens its buckets with the earlier pieces' sums)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm_multi.py tests/test_gpu_mgpu.py tests/test_gpu_msm_affine.py -x -q -m gpu > gpurun_out/r02_s2c8_tests.log 2>&1; tail -4 gpurun_out/r02_s2c8_tests.log
( for n in 21 22 23 24; do timeout 200 python tools/mgpu_bench.py $n 0 2>&1 | grep "e2e\|resident"; done
for cuts in "3,16" "6,24" "4,28" "8,32"; do
echo "== 2^24 linked, cuts=$cuts"; KZGB200_HOST_PIECES=$cuts timeout 200 python tools/mgpu_bench.py 24 0 2>&1 | grep e2e
done
echo "== 2^21 three pieces 4,20"; KZGB200_HOST_PIECES=4,20 timeout 200 python tools/mgpu_bench.py 21 0 2>&1 | grep e2e
timeout 200 python tools/msm_phases.py 20 21 2>&1 | grep msm ) 2>&1 | tee gpurun_out/r02_s2c8_carry.log
