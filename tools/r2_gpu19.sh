#!/bin/bash
# round 2, GPU call 19: direction rows through the texture pipe (experiment, ZCORDIC_TD_TEX=1): tests + A/B
mkdir -p gpurun_out
ZCORDIC_TD_TEX=1 timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "merged or nco or other_vectors or auto_selected or random_configurations" > gpurun_out/r2_pytest19.log 2>&1; echo "pytest(tex) rc=$?" >> gpurun_out/r2_pytest19.log
tail -3 gpurun_out/r2_pytest19.log
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz_min_under_load"], d["clocks"]["reasons"], d["parity_spot_check"])'
b() { timeout 300 python bench.py --no-cpu --no-e2e --no-configs --no-sustained "$@" 2>&1 | tail -1 | python -c "$fmt" "$*"; }
{
for t in 0 1 0 1; do
echo -n "TD_TEX=$t "; ZCORDIC_TD_TEX=$t b --steps 20 --warmup 3 --workload nco_cfg1
echo -n "TD_TEX=$t "; ZCORDIC_TD_TEX=$t b --steps 20 --warmup 3 --workload rotate_cfg1 --phase random
done
} > gpurun_out/r2_ab19.txt 2>&1
cat gpurun_out/r2_ab19.txt
