#!/bin/bash
# session 2, call 7: three linked pieces of the host-scalar MSM at 2^24 points (third arena): tests + cut sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm_multi.py tests/test_gpu_mgpu.py -x -q -m gpu > gpurun_out/r02_s2c7_tests.log 2>&1; tail -4 gpurun_out/r02_s2c7_tests.log
( echo "== 2^24 host_link=0 (two pieces, one reduction each)"; KZGB200_HOST_LINK=0 timeout 200 python tools/mgpu_bench.py 24 0 2>&1 | grep "e2e\|resident"
for cuts in "" "3,16" "4,24" "6,24" "2,12" "5,28" "12,64"; do
echo "== 2^24 linked, cuts=${cuts:-default 4,20}"; KZGB200_HOST_PIECES=$cuts timeout 200 python tools/mgpu_bench.py 24 0 2>&1 | grep e2e
done
for n in 22 23; do for cuts in "" "4,20"; do
echo "== 2^$n linked, cuts=${cuts:-default (two pieces 3/16)}"; KZGB200_HOST_PIECES=$cuts timeout 200 python tools/mgpu_bench.py $n 0 2>&1 | grep e2e
done; done ) 2>&1 | tee gpurun_out/r02_s2c7_pieces.log
