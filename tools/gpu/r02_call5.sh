#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py -x -q -m gpu -k "ntt or extend or polynomial or grand" > gpurun_out/r02_c5_tests.log 2>&1; tail -3 gpurun_out/r02_c5_tests.log
timeout 900 python -m pytest tests/test_gpu_large_parity.py -x -q -m gpu -k "ntt" >> gpurun_out/r02_c5_tests.log 2>&1; tail -3 gpurun_out/r02_c5_tests.log
timeout 900 python -m pytest tests/test_gpu_prover.py -x -q -m gpu >> gpurun_out/r02_c5_tests.log 2>&1; tail -3 gpurun_out/r02_c5_tests.log
for n in 16 20 22 24; do timeout 100 python tools/ntt_once.py $n; done 2>&1 | tee gpurun_out/r02_c5_ntt.log
timeout 300 python tools/prove_once.py 20 gs 4 2>&1 | tee gpurun_out/r02_c5_prove20.log
