#!/bin/bash
# tools/bench_all.sh -- every workload of bench.py once (sweep; plus random for the table-driven ones); under gpurun.
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s", round(d["roofline"]["achieved"]), "GB/s frac", round(d["roofline"]["frac"],3), "sm", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["parity_spot_check"])'
for w in rotate_cfg1 rotate_cfg1_noseed rotate_xy_cfg1 topolar_cfg2 nco_cfg1 sintable_p17 quarterwav_p18 quadtbl_p18; do
  python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w 2>&1 | tail -1 | python -c "$fmt" "$w sweep"
done
for w in rotate_cfg1 sintable_p17 quarterwav_p18; do
  python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w --phase random 2>&1 | tail -1 | python -c "$fmt" "$w random"
done
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload rotate_cfg1 --phase random --seed-mode packed 2>&1 | tail -1 | python -c "$fmt" "rotate_cfg1 random packed"
