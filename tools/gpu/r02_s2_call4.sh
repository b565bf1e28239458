#!/bin/bash
# session 2, call 4: a batched-affine round as one persistent kernel (forward / in-block inversion / backward)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_msm_affine.py -x -q -m gpu ) > gpurun_out/r02_s2c4_tests.log 2>&1; tail -6 gpurun_out/r02_s2c4_tests.log
for f in 1 0; do echo "== aff_fused=$f"; KZGB200_AFF_FUSED=$f timeout 300 python tools/msm_phases.py 20 21 22 24 2>&1 | grep msm; done | tee gpurun_out/r02_s2c4_phases.log
for r in 1 2; do echo "== 2^20 fused rounds=$r"; ROUNDS=$r timeout 100 python tools/msm_phases.py 20 2>&1 | grep msm; done | tee -a gpurun_out/r02_s2c4_phases.log
for r in 3; do echo "== 2^21 fused rounds=$r"; ROUNDS=$r timeout 100 python tools/msm_phases.py 21 2>&1 | grep msm; done | tee -a gpurun_out/r02_s2c4_phases.log
