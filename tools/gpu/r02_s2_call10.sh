#!/bin/bash
# session 2, call 10: backward pass of the affine rounds driven by the forward pass's operand descriptors (no cursor)
mkdir -p gpurun_out
for d in 1 2; do KZGB200_AFF_DESC=$d timeout 600 python -m pytest tests/test_gpu_msm_affine.py -x -q -m gpu 2>&1 | tail -1; done
( for d in 0 1 2; do echo "== aff_desc=$d"; KZGB200_AFF_DESC=$d timeout 300 python tools/msm_phases.py 22 24 2>&1 | grep msm; done ) | tee gpurun_out/r02_s2c10_desc.log
