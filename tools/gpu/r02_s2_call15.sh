#!/bin/bash
# session 2, call 15: cuts of the three linked pieces and chunks per round, with the final kernels
mkdir -p gpurun_out
( for cuts in "" "2,14" "3,20" "4,16" "4,20" "6,22"; do
echo "== 2^24 cuts=${cuts:-default 3,16}"; KZGB200_HOST_PIECES=$cuts timeout 200 python tools/mgpu_bench.py 24 0 2>&1 | grep e2e
done
for cuts in "" "12,64" "8,64" "4,20"; do
echo "== 2^21 cuts=${cuts:-default (two pieces 12,64)}"; KZGB200_HOST_PIECES=$cuts timeout 200 python tools/mgpu_bench.py 21 0 2>&1 | grep e2e
done
CHUNKS=1,2,3 timeout 300 python tools/msm_phases.py 21 24 2>&1 | grep msm ) 2>&1 | tee gpurun_out/r02_s2c15_cuts.log
