#!/bin/bash
# round 2, GPU call 3: comb with transposed 256-bit stores (tests + A/B), FP32 vectoring ubench, host chunk sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_runtime.py -m gpu -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
tail -6 gpurun_out/r2_pytest3.log
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz_min_under_load"], d["clocks"]["reasons"], d["parity_spot_check"])'
b() { timeout 300 python bench.py --no-cpu --no-e2e --no-configs --no-sustained "$@" 2>&1 | tail -1 | python -c "$fmt" "$*"; }
{
b --steps 20 --warmup 3 --workload nco_cfg1
b --steps 20 --warmup 3 --workload nco_cfg1 --no-comb
b --steps 20 --warmup 3 --workload nco_cfg1 --nco-step 0x9E3779B9
b --steps 20 --warmup 3 --workload nco_cfg1 --nco-step 0x00300000
b --steps 20 --warmup 3 --workload nco_cfg1 --nco-step 0x80000001
b --steps 20 --warmup 3 --workload nco_cfg1 --nco-step 0xDEADBEEF
b --steps 20 --warmup 3 --workload nco_cfg1 --nco-step 0x00012345
for K in 900 4500; do echo -n "K=$K "; ZCORDIC_COMB_K=$K b --steps 10 --warmup 3 --workload nco_cfg1; done
} > gpurun_out/r2_ab3.txt 2>&1
cat gpurun_out/r2_ab3.txt
./tools/ubench4 > gpurun_out/r2_ubench4.txt 2>&1; cat gpurun_out/r2_ubench4.txt
bash tools/ncu_capture.sh comb_cfg4 --workload nco_cfg1 > /dev/null 2>&1
rm -f gpurun_out/*.ncu-rep
grep -v "^$" gpurun_out/prof_comb_cfg4.md | head -50
e2e='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], d["e2e"])'
for lg in 20 22 24 26; do ZCORDIC_HOST_CHUNK_LG2=$lg timeout 300 python bench.py --no-cpu --no-configs --no-sustained --steps 5 --e2e-steps 3 2>&1 | tail -1 | python -c "$e2e" "chunk_lg2=$lg"; done > gpurun_out/r2_e2e_chunks.txt 2>&1
cat gpurun_out/r2_e2e_chunks.txt
python tools/pcie_probe.py > gpurun_out/r2_pcie_probe_n1.txt 2>&1; cat gpurun_out/r2_pcie_probe_n1.txt
