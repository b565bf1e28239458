#!/bin/bash
# final single-GPU records: ncu counts for traffic.json, bench (both arms)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fmaheavy.sum,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_msm24_launches.csv python tools/msm_once.py 24 0 2 > gpurun_out/r02_f1_ncu24.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:ntt_ --csv --log-file gpurun_out/r02_ntt24_launches.csv python tools/ntt_once.py 24 1 > gpurun_out/r02_f1_ncu_ntt.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_msm21_launches.csv python tools/msm_once.py 21 0 2 > gpurun_out/r02_f1_ncu21.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_prove20_launches.csv python tools/prove_once.py 20 gs 2 > gpurun_out/r02_f1_ncu_prove.log 2>&1
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/r02_f1_bench_ref.json 2> gpurun_out/r02_f1_bench_ref.err
tail -c 600 gpurun_out/r02_f1_bench_ref.json; tail -4 gpurun_out/r02_f1_bench_ref.err
echo done
