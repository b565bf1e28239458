#!/bin/bash
# session 2, call 9: the rounds rule again, now that the XYZZ walk is 8 % cheaper (1160 instead of 1360 wide MACs per entry)
mkdir -p gpurun_out
( ROUNDS=0,1,2,3 timeout 300 python tools/msm_phases.py 21 22 2>&1 | grep msm
ROUNDS=2,3,4 timeout 300 python tools/msm_phases.py 23 24 2>&1 | grep msm
ROUNDS=0,1 timeout 300 python tools/msm_phases.py 20 2>&1 | grep msm ) | tee gpurun_out/r02_s2c9_rounds.log
