#!/bin/bash
# tools/ncu_capture.sh <tag> [bench args...] -- run on the GPU box (under gpurun):
#   1. launch list of the bench command (gpu__time_duration per launch, cold-cache, serialised)
#   2. one --set full capture of the dominant kernel (k_rotate*), with source
# Results land in gpurun_out/ ; copy the summaries to profiles/ after reading them.
set -u
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-configs --no-sustained "$@" > gpurun_out/launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNEL:-k_rotate|k_topolar|k_lut|k_quadtbl}" -s ${NCU_SKIP:-3} -c 1 \
    -f -o gpurun_out/prof_${tag} python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs --no-sustained "$@" > gpurun_out/prof_${tag}.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_${tag}_raw.csv "ncu --set full, workload ${tag} ($*)" > gpurun_out/prof_${tag}.md 2>/dev/null
ls -la gpurun_out | tail -8
