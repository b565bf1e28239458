#!/bin/bash
# session 2, call 2: dedicated squaring (fp_sqr, 100 wide MACs): selftest, suite, phase times
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_s2c2_smoke.log 2>&1; tail -3 gpurun_out/r02_s2c2_smoke.log
( time timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_large_parity.py ) > gpurun_out/r02_s2c2_tests.log 2>&1; tail -6 gpurun_out/r02_s2c2_tests.log
timeout 300 python tools/msm_phases.py 20 21 22 24 2>&1 | grep msm | tee gpurun_out/r02_s2c2_phases.log
timeout 300 python tools/prove_once.py 20 gs 4 2>&1 | tail -4 | tee gpurun_out/r02_s2c2_prove20.log
