#!/bin/bash
# round 2, GPU call 6: per-sample rotation with the word-table suffix (tests + A/B), full gpu tier
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
tail -8 gpurun_out/r2_pytest6.log
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz_min_under_load"], d["clocks"]["reasons"], d["parity_spot_check"])'
b() { timeout 300 python bench.py --no-cpu --no-e2e --no-configs --no-sustained "$@" 2>&1 | tail -1 | python -c "$fmt" "$*"; }
{
b --steps 20 --warmup 3 --workload rotate_xy_cfg1
b --steps 20 --warmup 3 --workload rotate_xy_cfg1 --no-dp2a
b --steps 20 --warmup 3 --workload rotate_xy_cfg1 --phase random
b --steps 100 --warmup 3 --workload rotate_xy_cfg1
} > gpurun_out/r2_ab6.txt 2>&1
cat gpurun_out/r2_ab6.txt
NCU_KERNEL=k_rotate_dirs NCU_SKIP=3 bash tools/ncu_capture.sh rotxy_words --workload rotate_xy_cfg1 > /dev/null 2>&1
rm -f gpurun_out/*.ncu-rep
grep -v "^$" gpurun_out/prof_rotxy_words.md | head -45
