#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fmaheavy.sum,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:ntt_ --csv --log-file gpurun_out/r02_ntt24_launches.csv python tools/ntt_once.py 24 1 > gpurun_out/r02_c12_ncu_ntt.log 2>&1
grep -c ntt_ gpurun_out/r02_ntt24_launches.csv
