#!/bin/bash
mkdir -p gpurun_out
for il in 2 1 0; do KZGB200_AFF_INTERLEAVE=$il timeout 600 python -m pytest tests/test_gpu_msm_affine.py -x -q -m gpu > gpurun_out/r02_s2c13_tests_il$il.log 2>&1; tail -1 gpurun_out/r02_s2c13_tests_il$il.log; done
timeout 600 python -m pytest tests/test_gpu_msm_multi.py tests/test_gpu_prover.py -x -q -m gpu 2>&1 | tail -1
timeout 300 python tools/msm_phases.py 20 21 22 23 24 2>&1 | grep msm | tee gpurun_out/r02_s2c13_phases.log
ROUNDS=0,1,2 timeout 300 python tools/msm_phases.py 21 2>&1 | grep msm | tee -a gpurun_out/r02_s2c13_phases.log
