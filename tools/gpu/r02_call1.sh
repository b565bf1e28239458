#!/bin/bash
# round 2, GPU call 1: the missing gather working-set curve, baseline phase times, launch lists of the small-shard MSM
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/r02_c1_smi.log 2>&1
timeout 300 tools/micro/bin/gather sweep > gpurun_out/r02_gather_sweep.log 2>&1
timeout 300 python tools/msm_phases.py 20 21 22 23 24 > gpurun_out/r02_c1_phases_base.log 2>&1
KZGB200_AFF_ROUNDS=3 timeout 200 python tools/msm_phases.py 21 > gpurun_out/r02_c1_phases_21_r3.log 2>&1
KZGB200_AFF_ROUNDS=2 timeout 200 python tools/msm_phases.py 20 21 > gpurun_out/r02_c1_phases_r2.log 2>&1
KZGB200_AFF_ROUNDS=3 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_c1_msm21_r3_launches.csv python tools/msm_once.py 21 0 2 > gpurun_out/r02_c1_ncu21.log 2>&1
timeout 300 python tools/prove_once.py 20 gs 4 > gpurun_out/r02_c1_prove20.log 2>&1
echo done
