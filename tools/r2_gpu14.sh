#!/bin/bash
# round 2, GPU call 14: bench.py with every workload as the headline (e2e through the matching zc_*_host call)
mkdir -p gpurun_out
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s frac", round(d["roofline"]["frac"],3), "e2e", d["e2e"] and round(d["e2e"]["value"],2), d["parity_spot_check"])'
for w in nco_cfg1 topolar_cfg2 rotate_xy_cfg1 sintable_p17 sintable_p23 quarterwav_p25 quadtbl_p18 topolar_i16_cfg2 rotate_o16_cfg0 rotate_cfg1_noseed; do
  timeout 300 python bench.py --no-cpu --no-configs --no-sustained --steps 10 --warmup 3 --workload $w 2>&1 | tail -1 | python -c "$fmt" "$w" 2>&1 | tail -1
done > gpurun_out/r2_bench_each_workload.txt
cat gpurun_out/r2_bench_each_workload.txt
