#!/bin/bash
# Rebuild the tracked summaries under profiles/ from the ncu CSVs a GPU call left in gpurun_out/
# (tools/gpu/r02_s2_finalA.sh writes them).  Run from the repo root after the sources are final: traffic.json is keyed by a
# hash of the kernel sources and bench.py ignores entries whose hash no longer matches.
set -e
S=tools/summarize_profiles.py
ACC='msm_aff|batch_|msm_accumulate|msm_offsets|msm_scan'
python $S traffic gpurun_out/r02_msm24_launches.csv msm_accumulation_2_24 msm_part_hist "$ACC" msm \
  "one 2^24-point SRS MSM, table c = 22, 4 batched-affine rounds in 2 chunks: affine rounds + XYZZ walk (+ their scans)"
python $S traffic gpurun_out/r02_ntt24_launches.csv ntt_2_24 ntt_strided 'ntt_' ntt \
  "one forward 2^24 NTT: 2 strided passes + last pass (direct twiddle table at the first boundary)" 3
python $S multi gpurun_out/r02_msm24_launches.csv profiles/r02_msm24_launches.md \
  "round 2 — every launch of one 2^24-point SRS MSM (table c = 22, 4 batched-affine rounds in 2 chunks each)" msm_part_hist
python $S multi gpurun_out/r02_msm21_launches.csv profiles/r02_msm21_launches.md \
  "round 2 — every launch of one 2^21-point SRS MSM (the 8-GPU shard: table c = 20, 2 batched-affine rounds)" msm_part_hist
python $S multi gpurun_out/r02_ntt24_launches.csv profiles/r02_ntt24_launches.md \
  "round 2 — the passes of a 2^24 NTT (forward, then inverse)" ntt_strided
python $S launches gpurun_out/r02_prove20_launches.csv profiles/r02_prove20_launches.md "round 2 — launch list of one grand-sum proof at n = 2^20" 203
