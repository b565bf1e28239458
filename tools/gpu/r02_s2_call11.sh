#!/bin/bash
# session 2, call 11: descriptors as the only backward path: tests, then outputs-per-thread and chunk sweeps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm_affine.py tests/test_gpu_msm_multi.py -x -q -m gpu 2>&1 | tail -1
( for m in 8 12 16 24 32; do echo "== aff_m=$m"; KZGB200_AFF_M=$m timeout 300 python tools/msm_phases.py 22 24 2>&1 | grep msm; done
for c in 1 3 4; do echo "== aff_chunks=$c"; KZGB200_AFF_CHUNKS=$c timeout 300 python tools/msm_phases.py 24 2>&1 | grep msm; done ) | tee gpurun_out/r02_s2c11_m.log
