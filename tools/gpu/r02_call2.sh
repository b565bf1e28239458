#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_msm_multi.py tests/test_gpu_msm_affine.py tests/test_gpu_primitives.py tests/test_gpu_prover.py -x -q -m gpu > gpurun_out/r02_c2_tests.log 2>&1
tail -5 gpurun_out/r02_c2_tests.log
timeout 300 python tools/msm_phases.py 20 21 22 24 > gpurun_out/r02_c2_phases.log 2>&1
cat gpurun_out/r02_c2_phases.log
for ch in 1 2 4 8; do KZGB200_AFF_CHUNKS=$ch KZGB200_AFF_ROUNDS=2 timeout 200 python tools/msm_phases.py 20 21 >> gpurun_out/r02_c2_phases_chunks.log 2>&1; done
KZGB200_AFF_ROUNDS=3 timeout 200 python tools/msm_phases.py 21 >> gpurun_out/r02_c2_phases_chunks.log 2>&1
cat gpurun_out/r02_c2_phases_chunks.log
timeout 300 python tools/prove_once.py 20 gs 4 > gpurun_out/r02_c2_prove20.log 2>&1
KZGB200_MSM_MERGE=0 timeout 300 python tools/prove_once.py 20 gs 4 > gpurun_out/r02_c2_prove20_nomerge.log 2>&1
tail -2 gpurun_out/r02_c2_prove20.log gpurun_out/r02_c2_prove20_nomerge.log
