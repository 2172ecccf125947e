#!/bin/bash
# round 2, GPU call 10: merged records under the word table (experiment), TS riding in the TP rows of the per-sample kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest10.log
tail -4 gpurun_out/r2_pytest10.log
ZCORDIC_WORDS_MERGE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_rtl_sweeps.py -m gpu -q -x -k "rotate_const or nco or rtl" > gpurun_out/r2_pytest10m.log 2>&1; echo "pytest(words merged) rc=$?" >> gpurun_out/r2_pytest10m.log
tail -3 gpurun_out/r2_pytest10m.log
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz_min_under_load"], d["clocks"]["reasons"], d["parity_spot_check"])'
b() { timeout 300 python bench.py --no-cpu --no-e2e --no-configs --no-sustained "$@" 2>&1 | tail -1 | python -c "$fmt" "$*"; }
{
for m in 0 1 0 1; do
echo -n "WORDS_MERGE=$m "; ZCORDIC_WORDS_MERGE=$m b --steps 20 --warmup 3 --workload rotate_cfg1
echo -n "WORDS_MERGE=$m "; ZCORDIC_WORDS_MERGE=$m b --steps 400 --warmup 3 --workload rotate_cfg1
done
echo -n "WORDS_MERGE=0 "; ZCORDIC_WORDS_MERGE=0 b --steps 20 --warmup 3 --workload nco_cfg1 --nco-step 0x100
echo -n "WORDS_MERGE=1 "; ZCORDIC_WORDS_MERGE=1 b --steps 20 --warmup 3 --workload nco_cfg1 --nco-step 0x100
echo -n "WORDS_MERGE=0 "; ZCORDIC_WORDS_MERGE=0 b --steps 20 --warmup 3 --workload nco_cfg1 --nco-step 0x80000001
echo -n "WORDS_MERGE=1 "; ZCORDIC_WORDS_MERGE=1 b --steps 20 --warmup 3 --workload nco_cfg1 --nco-step 0x80000001
echo -n "WORDS_MERGE=0 "; ZCORDIC_WORDS_MERGE=0 b --steps 20 --warmup 3 --workload rotate_o16_cfg0
echo -n "WORDS_MERGE=1 "; ZCORDIC_WORDS_MERGE=1 b --steps 20 --warmup 3 --workload rotate_o16_cfg0
b --steps 20 --warmup 3 --workload rotate_xy_cfg1
b --steps 20 --warmup 3 --workload rotate_xy_cfg1 --phase random
} > gpurun_out/r2_ab10.txt 2>&1
cat gpurun_out/r2_ab10.txt
