#!/bin/bash
# the whole -m gpu suite (minus the 2^22 proofs) once under the bounds-checking debug build
export KZGB200_DEBUG_BOUNDS=1
mkdir -p gpurun_out
( time python -m kzg_grandsums_study_b200.build ) > gpurun_out/r02_bounds_build.log 2>&1; tail -3 gpurun_out/r02_bounds_build.log
nm -D kzg_grandsums_study_b200/libkzgb200.so | grep -c g_dbg
( time timeout 3000 python -m pytest tests -q -m gpu --deselect "tests/test_gpu_large_parity.py::test_full_size_proofs_byte_identical_to_c_oracle" -k "not mgpu and not two_contexts" ) > gpurun_out/r02_bounds_tests.log 2>&1
tail -8 gpurun_out/r02_bounds_tests.log
