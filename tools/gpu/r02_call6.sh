#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ops_bench.py 16 20 22 24 2>&1 | tee gpurun_out/r02_c6_ops.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:ntt_ --csv --log-file gpurun_out/r02_c6_ntt24.csv python tools/ntt_once.py 24 1 > gpurun_out/r02_c6_ncu.log 2>&1
tail -2 gpurun_out/r02_c6_ncu.log
