#!/bin/bash
# Round measurement suite (run on the GPU box through gpurun): tests, bench (both arms), launch list, ncu capture.
# usage: bash tools/gpu_measure.sh <tag>      outputs under gpurun_out/
tag=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${tag}_pytest_gpu.log
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; tail -c 300 gpurun_out/${tag}_bench_n1.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref_n1.json 2> gpurun_out/${tag}_bench_ref.err; tail -c 200 gpurun_out/${tag}_bench_ref_n1.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --search-c4-tracks 0 > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k0_resample|sumsq|normalize|k1_spectral|group_|k2_fingerprint|compact" -c 16 \
    -o gpurun_out/${tag}_prof_index python tools/gpu_time_index.py 7200 2 1 > gpurun_out/${tag}_ncu_full.log 2>&1
tail -2 gpurun_out/${tag}_ncu_full.log
# search kernels in the config #4 regime (10 k tracks from PCM, 600 queries so that the ncu replays stay short)
ncu --set full --clock-control none --import-source on -k regex:"search_probe|search_select" -c 2 -f \
    -o gpurun_out/${tag}_prof_search python tools/bench_c4_pcm.py --queries 600 --ref-queries 0 --steps 1 > gpurun_out/${tag}_ncu_search.log 2>&1
tail -1 gpurun_out/${tag}_ncu_search.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
