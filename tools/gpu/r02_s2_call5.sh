#!/bin/bash
# session 2, call 5: where the fused round loses its time (modes 2 / 3 give wrong points: timing only)
mkdir -p gpurun_out
for f in 1 2 3; do echo "== aff_fused mode=$f (1: full, 2: no root inversion, 3: no tree at all)"; KZGB200_AFF_FUSED=$f timeout 300 python tools/msm_phases.py 21 24 2>&1 | grep msm; done | tee gpurun_out/r02_s2c5_modes.log
