#!/bin/bash
# Final profiling pass of a round, run on the GPU box by gpurun (one call).  Every command that runs under ncu has run
# plainly, with the same arguments, directly before (B200_PROFILING.md).  Outputs land in gpurun_out/ with the round's tag.
TAG=${1:-r02}
OUT=gpurun_out
set -x
# 1. launch list of the bench command (index measurement only), cold-cache per-launch times
python bench.py --steps 2 --warmup 1 --index-only > $OUT/${TAG}_bench_steps2.json 2> $OUT/${TAG}_bench_steps2.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --index-only > $OUT/${TAG}_ncu_launches.log 2>&1
# 2. the index kernels of configs[1], exact mode: one full capture of the second run's launches
python tools/profile_index.py --runs 2 > $OUT/${TAG}_profile_index.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k0_resample|sumsq|normalize|k1_|group_|k2_|compact' -s 15 -c 15 \
    -o $OUT/${TAG}_prof_index python tools/profile_index.py --runs 2 > $OUT/${TAG}_ncu_index.log 2>&1
# 3. the tolerance mode's K1
python tools/profile_index.py --runs 2 --fast > $OUT/${TAG}_profile_index_fast.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k1_fast -s 1 -c 1 \
    -o $OUT/${TAG}_prof_k1_fast python tools/profile_index.py --runs 2 --fast > $OUT/${TAG}_ncu_k1_fast.log 2>&1
# 4. the probe and ranking kernels on configs[3] (first 2000 of the 10 000 queries, the whole 10 000-track database)
python tools/bench_c4_pcm.py --ref-queries 0 --queries 2000 > $OUT/${TAG}_c4_q2000.json 2> $OUT/${TAG}_c4_q2000.err &&
ncu --set full --clock-control none --import-source on -k 'regex:search_pair|search_probe|search_select|select_merge' -s 4 -c 4 \
    -o $OUT/${TAG}_prof_search python tools/bench_c4_pcm.py --ref-queries 0 --queries 2000 > $OUT/${TAG}_ncu_search.log 2>&1
ls -la $OUT/${TAG}_prof_*.ncu-rep
