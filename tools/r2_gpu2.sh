#!/bin/bash
# round 2, GPU call 2: full gpu test tier (no -x), comb K sweep, ncu of a slow and a fast comb launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -8 gpurun_out/r2_pytest.log
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz_min_under_load"], d["clocks"]["reasons"], d["parity_spot_check"])'
b() { timeout 300 python bench.py --no-cpu --no-e2e --no-configs --no-sustained "$@" 2>&1 | tail -1 | python -c "$fmt" "$*"; }
{
for K in 4096 16384 65536 262144 1048576 4194304 16777216 67108864; do
  echo -n "K=$K "; ZCORDIC_COMB_K=$K b --steps 10 --warmup 3 --workload nco_cfg1 --nco-step 0x00300000
done
for K in 450 900 3600 28800 460800 7372800; do
  echo -n "K=$K "; ZCORDIC_COMB_K=$K b --steps 10 --warmup 3 --workload nco_cfg1
done
} > gpurun_out/r2_comb_ksweep.txt 2>&1
cat gpurun_out/r2_comb_ksweep.txt
ZCORDIC_COMB_K=4096 bash tools/ncu_capture.sh comb_k4096 --workload nco_cfg1 --nco-step 0x00300000 > /dev/null 2>&1
ZCORDIC_COMB_K=4194304 bash tools/ncu_capture.sh comb_k4m --workload nco_cfg1 --nco-step 0x00300000 > /dev/null 2>&1
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/prof_comb_k4096.md gpurun_out/prof_comb_k4m.md
