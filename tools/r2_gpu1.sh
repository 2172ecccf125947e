#!/bin/bash
# round 2, GPU call 1: gpu test tier, the default bench line (headline + configs + sustained), NCO comb A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
nproc >> gpurun_out/r2_gpu.txt; free -g >> gpurun_out/r2_gpu.txt; lscpu | grep -i "numa\|model name\|socket" >> gpurun_out/r2_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz_min_under_load"], d["clocks"]["reasons"], d["parity_spot_check"])'
b() { timeout 300 python bench.py --no-cpu --no-e2e --no-configs --no-sustained "$@" 2>&1 | tail -1 | python -c "$fmt" "$*"; }
{
b --steps 20 --warmup 3 --workload nco_cfg1
b --steps 20 --warmup 3 --workload nco_cfg1 --no-comb
b --steps 20 --warmup 3 --workload nco_cfg1 --nco-step 0x100
b --steps 20 --warmup 3 --workload nco_cfg1 --nco-step 0x9E3779B9
b --steps 20 --warmup 3 --workload nco_cfg1 --nco-step 0x00300000
b --steps 20 --warmup 3 --workload nco_cfg1 --nco-step 0x80000001
b --steps 20 --warmup 3 --workload rotate_cfg1
b --steps 20 --warmup 3 --workload rotate_cfg1 --phase random
} > gpurun_out/r2_ab1.txt 2>&1
cat gpurun_out/r2_ab1.txt
