#!/bin/bash
mkdir -p gpurun_out
nproc > gpurun_out/r02_c4_nproc.log
( time timeout 2400 python -m pytest tests/test_gpu_large_parity.py -x -q -m gpu --durations=20 ) > gpurun_out/r02_c4_large.log 2>&1
tail -30 gpurun_out/r02_c4_large.log
