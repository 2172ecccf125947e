#!/bin/bash
# tools/run_exp.sh -- A/B of experiment builds (exp/*.so, cordic_b200/build.py --out ... -D...) against the product library
#   EXTRA="--workload topolar_cfg2" STEPS="10 50" bash tools/run_exp.sh
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s frac", round(d["roofline"]["frac"],3), "ms", round(d["ms_per_step"],4), "sm", d["clocks"]["sm_mhz_min_under_load"], d["clocks"]["reasons"], d["parity_spot_check"])'
mkdir -p gpurun_out
for rep in 1 2; do
for st in ${STEPS:-20}; do
for lib in "" $(ls exp/*.so); do
  ZCORDIC_LIB=$lib python bench.py --steps $st --warmup 3 --no-cpu --no-e2e --no-configs --no-sustained --seed-mode words $EXTRA 2>&1 | tail -1 | python -c "$fmt" "$st steps lib=${lib:-product}"
done
done
done | tee gpurun_out/exp.txt
