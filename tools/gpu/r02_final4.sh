#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lagrange.py tests/test_gpu_mgpu.py -x -q -m gpu 2>&1 | tail -2
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fmaheavy.sum,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_msm24_launches.csv python tools/msm_once.py 24 0 2 > gpurun_out/r02_f4_ncu24.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_prove20_launches.csv python tools/prove_once.py 20 gs 2 > gpurun_out/r02_f4_ncu_prove.log 2>&1
echo done
