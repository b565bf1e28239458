#!/bin/bash
for v in 1 0 1 0; do echo "prefetch=$v"; KZGB200_NTT_PREFETCH=$v python tools/ops_bench.py 21 22 24 2>&1 | cut -c1-62; done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for v in 1 0; do KZGB200_NTT_PREFETCH=$v ncu --metrics $M --clock-control none -k regex:ntt_strided -c 1 --csv python tools/ntt_once.py 24 1 2>&1 | grep -E "dram__bytes_read|gpu__time" | cut -d, -f 13-15; done
