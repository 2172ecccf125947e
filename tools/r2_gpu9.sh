#!/bin/bash
# round 2, GPU call 9: merged (x, y, TS) records for the byte-table kernel -- tests + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest9.log
tail -6 gpurun_out/r2_pytest9.log
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz_min_under_load"], d["clocks"]["reasons"], d["parity_spot_check"])'
b() { timeout 300 python bench.py --no-cpu --no-e2e --no-configs --no-sustained "$@" 2>&1 | tail -1 | python -c "$fmt" "$*"; }
{
b --steps 20 --warmup 3 --workload rotate_cfg1 --phase random
b --steps 20 --warmup 3 --workload rotate_cfg1 --phase random --no-merge
b --steps 20 --warmup 3 --workload nco_cfg1
b --steps 20 --warmup 3 --workload nco_cfg1 --no-merge
b --steps 20 --warmup 3 --workload rotate_cfg1 --seed-mode packed
b --steps 20 --warmup 3 --workload rotate_cfg1 --seed-mode packed --no-merge
b --steps 20 --warmup 3 --workload rotate_cfg1
} > gpurun_out/r2_ab9.txt 2>&1
cat gpurun_out/r2_ab9.txt
