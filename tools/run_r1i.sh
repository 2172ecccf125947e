mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"],1), "GS/s frac", round(d["roofline"]["frac"],3), "ms", round(d["ms_per_step"],4), d["clocks"]["reasons"])'
b() { python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e "$@" 2>&1 | tail -1 | python -c "$fmt" "$*"; }
for w in sintable_p17 quarterwav_p18; do for ph in sweep random; do
  b --workload $w --phase $ph
done; done
