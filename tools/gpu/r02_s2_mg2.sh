#!/bin/bash
# session 2: the multi-device tests and the 2-GPU bench line on the final sources
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mgpu.py tests/test_gpu_msm_multi.py -x -q -m gpu 2>&1 | tail -2
bash tools/gpu/r02_s2_scale.sh 2
