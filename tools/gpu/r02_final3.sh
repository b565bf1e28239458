#!/bin/bash
mkdir -p gpurun_out
( time timeout 3000 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_large_parity.py::test_full_size_proofs_byte_identical_to_c_oracle ) > gpurun_out/r02_f3_tests.log 2>&1; tail -4 gpurun_out/r02_f3_tests.log
timeout 900 python -m pytest tests/test_gpu_large_parity.py -x -q -m gpu -k "plain-20 or selected_k2-20" 2>&1 | tail -2
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fmaheavy.sum,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 600 ncu --metrics $M --clock-control none -k regex:ntt_ --csv --log-file gpurun_out/r02_ntt24_launches.csv python tools/ntt_once.py 24 1 > gpurun_out/r02_f3_ncu_ntt.log 2>&1
echo done
