mkdir -p gpurun_out
python -m pytest tests/test_rtl_sweeps.py tests/test_rtl_vectors.py -m gpu -x -q 2>&1 | tail -3
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s frac", round(d["roofline"]["frac"],3), "sm", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["parity_spot_check"])'
for rep in 1 2; do
python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --seed-mode words 2>&1 | tail -1 | python -c "$fmt" "words dp2a 20 steps"
python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --seed-mode words --no-dp2a 2>&1 | tail -1 | python -c "$fmt" "words no-dp2a 20 steps"
done
python bench.py --steps 100 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "$fmt" "auto 100 steps"
python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --workload nco_cfg1 --nco-step 0x100 2>&1 | tail -1 | python -c "$fmt" "nco step 0x100"
