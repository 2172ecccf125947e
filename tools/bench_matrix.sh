#!/bin/bash
# tools/bench_matrix.sh -- one line per (workload, phase pattern, seed mode); run under gpurun.
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s", round(d["roofline"]["achieved"]), "GB/s frac", round(d["roofline"]["frac"],3), "sm", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["parity_spot_check"])'
for mode in words packed regs; do
  for ph in sweep random; do
    python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-configs --no-sustained --phase $ph --seed-mode $mode 2>&1 | tail -1 | python -c "$fmt" "rotate_cfg1 $ph $mode"
  done
  python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-configs --no-sustained --workload nco_cfg1 --seed-mode $mode 2>&1 | tail -1 | python -c "$fmt" "nco_cfg1 step=0x01234567 $mode"
done
python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-configs --no-sustained --workload nco_cfg1 --nco-step 0x00010000 2>&1 | tail -1 | python -c "$fmt" "nco_cfg1 step=0x00010000 auto"
python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-configs --no-sustained --workload nco_cfg1 --nco-step 0x00000100 2>&1 | tail -1 | python -c "$fmt" "nco_cfg1 step=0x00000100 auto"
