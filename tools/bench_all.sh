#!/bin/bash
# tools/bench_all.sh -- every workload of bench.py once (sweep; random where the phase pattern matters) plus the A/B
# switches the design document quotes; under gpurun.
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s", round(d["roofline"]["achieved"]), "GB/s frac", round(d["roofline"]["frac"],3), "sm", d["clocks"]["sm_mhz_min_under_load"], d["clocks"]["reasons"], d["parity_spot_check"])'
b() { python bench.py --no-cpu --no-e2e --no-configs --no-sustained "$@" 2>&1 | tail -1 | python -c "$fmt" "$*"; }
for w in rotate_cfg1 rotate_cfg1_noseed rotate_xy_cfg1 topolar_cfg2 nco_cfg1 sintable_p17 quarterwav_p18 quadtbl_p18; do
  b --steps 10 --warmup 3 --workload $w
done
for w in rotate_cfg1 rotate_xy_cfg1 sintable_p17 quarterwav_p18 quadtbl_p18; do
  b --steps 10 --warmup 3 --workload $w --phase random
done
b --steps 10 --warmup 3 --phase random --seed-mode words
b --steps 10 --warmup 3 --workload topolar_cfg2 --no-tail
b --steps 10 --warmup 3 --seed-mode words --no-dp2a
b --steps 10 --warmup 3 --workload nco_cfg1 --nco-step 0x100
ZCORDIC_LUT_SMEM=0 b --steps 10 --warmup 3 --workload sintable_p17 --phase random
ZCORDIC_LUT_SMEM=0 b --steps 10 --warmup 3 --workload quarterwav_p18 --phase random
b --steps 20 --warmup 3 --seed-mode words
b --steps 100 --warmup 3 --seed-mode words
b --steps 400 --warmup 3 --seed-mode words
