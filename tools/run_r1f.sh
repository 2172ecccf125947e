mkdir -p gpurun_out
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s frac", round(d["roofline"]["frac"],3), "ms", round(d["ms_per_step"],4), "sm", d["clocks"]["sm_mhz"], d["clocks"]["sm_mhz_min_under_load"], d["clocks"]["power_w_max"], d["clocks"]["reasons"])'
for rep in 1 2 3; do
for st in 10 50; do
python bench.py --steps $st --warmup 3 --no-cpu --no-e2e --seed-mode words 2>&1 | tail -1 | python -c "$fmt" "dp2a $st steps"
python bench.py --steps $st --warmup 3 --no-cpu --no-e2e --seed-mode words --no-dp2a 2>&1 | tail -1 | python -c "$fmt" "no-dp2a $st steps"
done
done
ncu --set full --clock-control none --import-source on -k regex:'k_rotate_seeded' -s 3 -c 1 -f -o gpurun_out/prof_dp python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --seed-mode words > gpurun_out/prof_dp.log 2>&1
ncu -i gpurun_out/prof_dp.ncu-rep --page raw --csv > gpurun_out/prof_dp_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_dp_raw.csv "k_rotate_seeded<8,0,1,3> (IDP.2A suffix), cfg1 sweep" > gpurun_out/r1f_ncu_dp.md
cat gpurun_out/r1f_ncu_dp.md | cut -c1-140
rm -f gpurun_out/*.ncu-rep
