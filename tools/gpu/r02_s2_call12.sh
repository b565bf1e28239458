#!/bin/bash
# session 2, call 12: outputs of an affine round dealt to the lanes of a warp one by one (coalesced dense reads / writes)
mkdir -p gpurun_out
for il in 2 1; do KZGB200_AFF_INTERLEAVE=$il timeout 600 python -m pytest tests/test_gpu_msm_affine.py -x -q -m gpu 2>&1 | tail -1; done
( for il in 0 1 2; do echo "== aff_interleave=$il"; KZGB200_AFF_INTERLEAVE=$il timeout 300 python tools/msm_phases.py 22 24 2>&1 | grep msm; done
for m in 48 64; do echo "== aff_interleave=1 aff_m=$m"; KZGB200_AFF_INTERLEAVE=1 KZGB200_AFF_M=$m timeout 300 python tools/msm_phases.py 24 2>&1 | grep msm; done
echo "== 2^21 rounds 0 / 2 (interleave 1)"; KZGB200_AFF_INTERLEAVE=1 ROUNDS=0,2 timeout 300 python tools/msm_phases.py 21 2>&1 | grep msm ) | tee gpurun_out/r02_s2c12_il.log
