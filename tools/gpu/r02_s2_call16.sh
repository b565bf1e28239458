#!/bin/bash
# session 2, call 16: coordinates as one 256-bit access (LDG.E.256 / STG.E.256) in the accumulation kernels -- a variant
# build (kzg_grandsums_study_b200/variants/libkzgb200_v8.so, the patch of profiles/experiments) against the library: tests, A/B
mkdir -p gpurun_out
V=kzg_grandsums_study_b200/variants/libkzgb200_v8.so
KZGB200_LIB=$V timeout 900 python -m pytest tests/test_gpu_msm_affine.py tests/test_gpu_msm_multi.py tests/test_gpu_primitives.py -x -q -m gpu 2>&1 | tail -1
( for lib in "" $V; do echo "== lib: ${lib:-library}"; KZGB200_LIB=$lib timeout 300 python tools/msm_phases.py 20 21 22 24 2>&1 | grep msm; done ) | tee gpurun_out/r02_s2c16_v8.log
