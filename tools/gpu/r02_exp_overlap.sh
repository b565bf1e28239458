#!/bin/bash
for c1 in 2 4 8 16; do for sm in 0 60 72 100; do
echo "== CHUNKS1=$c1 B1SMEM=$sm"; CHUNKS1=$c1 B1SMEM=$sm timeout 200 python tools/msm_phases.py 24 2>&1 | grep msm
done; done 2>&1 | tee gpurun_out/r02_exp_overlap.log
for c1 in 4 8; do for sm in 0 72; do
echo "== 2^21 CHUNKS1=$c1 B1SMEM=$sm"; CHUNKS1=$c1 B1SMEM=$sm timeout 200 python tools/msm_phases.py 21 2>&1 | grep msm
done; done 2>&1 | tee -a gpurun_out/r02_exp_overlap.log
