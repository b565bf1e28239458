#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mgpu.py tests/test_gpu_msm_multi.py -x -q -m gpu > gpurun_out/r02_c9_tests.log 2>&1; tail -15 gpurun_out/r02_c9_tests.log
timeout 300 python tools/mgpu_bench.py 24 0 2>&1 | tee gpurun_out/r02_c9_mgpu1.log
timeout 300 python tools/mgpu_bench.py 22 0,0 2>&1 | tee -a gpurun_out/r02_c9_mgpu1.log
