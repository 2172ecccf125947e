# tools/run_final.sh -- under gpurun (1 GPU): the round's final evidence set
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest_gpu.log 2>&1; tail -4 gpurun_out/final_pytest_gpu.log
bash tools/bench_all.sh > gpurun_out/final_bench_all.txt 2>&1
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s", round(d["roofline"]["achieved"]), "GB/s frac", round(d["roofline"]["frac"],3), "sm", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["parity_spot_check"])'
b() { python bench.py --no-cpu --no-e2e --no-configs --no-sustained "$@" 2>&1 | tail -1 | python -c "$fmt" "$*" >> gpurun_out/final_bench_all.txt; }
b --steps 10 --warmup 3 --workload topolar_cfg2 --no-tail
b --steps 10 --warmup 3 --seed-mode words --no-dp2a
b --steps 10 --warmup 3 --workload nco_cfg1 --nco-step 0x100
b --steps 10 --warmup 3 --workload quadtbl_p18 --phase random
b --steps 20 --warmup 3 --seed-mode words
b --steps 100 --warmup 3 --seed-mode words
b --steps 400 --warmup 3 --seed-mode words
cat gpurun_out/final_bench_all.txt
bash tools/ncu_final.sh > gpurun_out/final_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_final_raw.csv "k_rotate_seeded<8,0,1,3> (word table, IDP.2A suffix), cfg1 sweep, 2^30 samples" > gpurun_out/final_ncu_seeded.md
ncu -i gpurun_out/prof_final.ncu-rep --page source --csv > gpurun_out/prof_final_source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
./tools/membench > gpurun_out/final_membench.txt 2>&1
tail -c 2600 gpurun_out/bench_final_n1.json; echo; cat gpurun_out/bench_final_ref.json
