mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "quadtbl or qtbl or table_cores" 2>&1 | tail -2
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s frac", round(d["roofline"]["frac"],3), "ms", round(d["ms_per_step"],4), d["clocks"]["reasons"], d["parity_spot_check"])'
b() { python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e "$@" 2>&1 | tail -1 | python -c "$fmt" "$*"; }
b --workload quadtbl_p18
b --workload quadtbl_p18 --phase random
