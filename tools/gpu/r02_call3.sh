#!/bin/bash
mkdir -p gpurun_out
python tools/debug/multi_dbg.py 2>&1 | grep -c "bad=\[\]" 
python tools/debug/multi_dbg.py 2>&1 | grep -v "bad=\[\]" | head
for ch in 1 2 4; do echo "== chunks $ch rounds 2"; KZGB200_AFF_CHUNKS=$ch KZGB200_AFF_ROUNDS=2 timeout 200 python tools/msm_phases.py 20 21; done 2>&1 | tee gpurun_out/r02_c3_chunks.log
for ch in 1 2 4; do echo "== chunks $ch default rounds"; KZGB200_AFF_CHUNKS=$ch timeout 200 python tools/msm_phases.py 22 24; done 2>&1 | tee -a gpurun_out/r02_c3_chunks.log
timeout 1500 python -m pytest tests/test_gpu_msm_multi.py tests/test_gpu_msm_affine.py tests/test_gpu_primitives.py tests/test_gpu_prover.py -x -q -m gpu > gpurun_out/r02_c3_tests.log 2>&1
tail -5 gpurun_out/r02_c3_tests.log
