#!/bin/bash
# round 2, GPU call 16: NCO through the LUT cores (tests + the new bench rows)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_runtime.py -m gpu -q -x -k "lut" > gpurun_out/r2_pytest16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest16.log
tail -4 gpurun_out/r2_pytest16.log
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz_min_under_load"], d["clocks"]["reasons"], d["parity_spot_check"])'
b() { timeout 300 python bench.py --no-cpu --no-e2e --no-configs --no-sustained "$@" 2>&1 | tail -1 | python -c "$fmt" "$*"; }
{
b --steps 20 --warmup 3 --workload nco_sintable_p17
b --steps 20 --warmup 3 --workload nco_sintable_p17 --nco-step 0x100
b --steps 20 --warmup 3 --workload nco_quarterwav_p18
b --steps 20 --warmup 3 --workload nco_quarterwav_p18 --nco-step 0x100
} > gpurun_out/r2_ab16.txt 2>&1
cat gpurun_out/r2_ab16.txt
