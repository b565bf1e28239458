#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02_f2_smoke.log 2>&1; tail -3 gpurun_out/r02_f2_smoke.log
( time timeout 3000 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02_f2_tests.log 2>&1; tail -6 gpurun_out/r02_f2_tests.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fmaheavy.sum,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 600 ncu --metrics $M --clock-control none -k regex:ntt_ --csv --log-file gpurun_out/r02_ntt24_launches.csv python tools/ntt_once.py 24 1 > gpurun_out/r02_f2_ncu_ntt.log 2>&1
echo done
