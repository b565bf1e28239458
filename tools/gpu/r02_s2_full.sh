#!/bin/bash
# session 2: one `ncu --set full` capture of the top kernels of the final sources (after the same command exited 0 without ncu)
mkdir -p gpurun_out
timeout 200 python tools/msm_once.py 24 0 1 > gpurun_out/r02_s2_full_plain.log 2>&1 || exit 1
tail -1 gpurun_out/r02_s2_full_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"msm_aff_backward|msm_aff_forward|msm_accumulate" -c 5 -f -o gpurun_out/r02_s2_msm_top python tools/msm_once.py 24 0 1 > gpurun_out/r02_s2_full_ncu.log 2>&1
tail -2 gpurun_out/r02_s2_full_ncu.log; ls -la gpurun_out/r02_s2_msm_top.ncu-rep
