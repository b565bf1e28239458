#!/bin/bash
# session 2, final pass A: the whole -m gpu suite, smoke, and the ncu count / launch-list passes of the final sources
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_s2fA_smoke.log 2>&1; tail -2 gpurun_out/r02_s2fA_smoke.log
( time timeout 3000 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02_s2fA_tests.log 2>&1; tail -6 gpurun_out/r02_s2fA_tests.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fmaheavy.sum,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_msm24_launches.csv python tools/msm_once.py 24 0 2 > gpurun_out/r02_s2fA_ncu24.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:ntt_ --csv --log-file gpurun_out/r02_ntt24_launches.csv python tools/ntt_once.py 24 1 > gpurun_out/r02_s2fA_ncu_ntt.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_msm21_launches.csv python tools/msm_once.py 21 0 2 > gpurun_out/r02_s2fA_ncu21.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_prove20_launches.csv python tools/prove_once.py 20 gs 2 > gpurun_out/r02_s2fA_ncu_prove.log 2>&1
tail -2 gpurun_out/r02_s2fA_ncu24.log gpurun_out/r02_s2fA_ncu_ntt.log gpurun_out/r02_s2fA_ncu21.log gpurun_out/r02_s2fA_ncu_prove.log
echo done
