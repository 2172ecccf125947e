#!/bin/bash
# round 2, GPU call 4: comb with half-trading for K = 2 (mod 4), ubench4 fixed, ncu of the cfg4 comb launch
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "nco" > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log
tail -6 gpurun_out/r2_pytest4.log
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz_min_under_load"], d["clocks"]["reasons"], d["parity_spot_check"])'
b() { timeout 300 python bench.py --no-cpu --no-e2e --no-configs --no-sustained "$@" 2>&1 | tail -1 | python -c "$fmt" "$*"; }
{
b --steps 20 --warmup 3 --workload nco_cfg1
b --steps 100 --warmup 3 --workload nco_cfg1
b --steps 20 --warmup 3 --workload nco_cfg1 --nco-step 0x00012345
b --steps 20 --warmup 3 --workload nco_cfg1 --nco-step 0x80000001
for K in 2700 6300; do echo -n "K=$K "; ZCORDIC_COMB_K=$K b --steps 10 --warmup 3 --workload nco_cfg1; done
} > gpurun_out/r2_ab4.txt 2>&1
cat gpurun_out/r2_ab4.txt
./tools/ubench4 > gpurun_out/r2_ubench4.txt 2>&1; cat gpurun_out/r2_ubench4.txt
NCU_KERNEL=k_rotate_seeded NCU_SKIP=3 bash tools/ncu_capture.sh comb_cfg4 --workload nco_cfg1 > /dev/null 2>&1
rm -f gpurun_out/*.ncu-rep
grep -v "^$" gpurun_out/prof_comb_cfg4.md | head -50
