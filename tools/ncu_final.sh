#!/bin/bash
# tools/ncu_final.sh -- run on the GPU box (under gpurun): the evidence bench.py's roofline object points at.
#   1. launch list of the DEFAULT bench command (auto flavour: probe + word-table kernel + gated-off byte-table kernel)
#   2. one --set full capture of the word-table kernel (same kernel, forced so that -s/-c pick it)
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
    --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-configs --no-sustained > gpurun_out/launches_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_rotate_seeded' -s 3 -c 1 \
    -f -o gpurun_out/prof_final python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs --no-sustained --seed-mode words > gpurun_out/prof_final.log 2>&1
ncu -i gpurun_out/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null
python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_final_ref.json 2>> gpurun_out/bench_final_n1.err
tail -c 1500 gpurun_out/bench_final_n1.json
