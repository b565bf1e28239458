#!/bin/bash
# session 2, call 14: the rounds rule once more (descriptors + interleaved outputs), prove times
mkdir -p gpurun_out
( ROUNDS=0,1,2 timeout 300 python tools/msm_phases.py 20 2>&1 | grep msm
ROUNDS=2,3,4 timeout 300 python tools/msm_phases.py 22 2>&1 | grep msm
ROUNDS=3,4 timeout 300 python tools/msm_phases.py 23 2>&1 | grep msm
ROUNDS=4,5 timeout 300 python tools/msm_phases.py 24 2>&1 | grep msm ) | tee gpurun_out/r02_s2c14_rounds.log
timeout 300 python tools/prove_once.py 20 gs 4 2>&1 | tail -3 | tee gpurun_out/r02_s2c14_prove20.log
KZGB200_AFF_MIN_ENTRIES_LOG=23 timeout 300 python tools/prove_once.py 20 gs 4 2>&1 | tail -2 | tee -a gpurun_out/r02_s2c14_prove20.log
