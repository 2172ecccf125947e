#!/bin/bash
# round 2, final state on 1 GPU: what the driver runs (gpu tier, smoke, default bench line) + the launch list of the bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_final_pytest.log; tail -3 gpurun_out/r2_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; echo "bench rc=$?"; tail -4 gpurun_out/r2_final_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench_n1.json').read().strip().splitlines()[-1])
print("value",d["value"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"], "launches",d["gpu_launches"], d["clocks"])
print("sustained", d["sustained"]["value"], d["sustained"]["roofline"]["frac"], d["sustained"]["clocks"]["sm_mhz"])
for c in d["configs"]:
    if "error" in c: print(c); continue
    print("%-20s %-8s %7.1f GS/s frac %.3f steps %d ok=%s clk=%s %s e2e=%s" % (c["workload"], c["phase"][:8], c["value"], c["roofline"]["frac"], c["steps"], c["parity_spot_check"], c["clocks"]["sm_mhz"], c["clocks"]["reasons"], (c.get("e2e") or {}).get("value")))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_default.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --sustained-steps 5 --configs-seconds 0 > gpurun_out/r2_launches_default.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2_launches_default.csv')) if len(r) > 5 and r[0].strip('"').isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4]; val = float(r[-1].replace(',', '')); unit = r[-2]
    a = agg.setdefault(name[:110], [0, 0.0, unit]); a[0] += 1; a[1] += val
with open('gpurun_out/r2_launches_default.txt', 'w') as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --sustained-steps 5 --configs-seconds 0\n")
    f.write("# kernel | launches | total duration | unit (cold-cache, serialised; torch's at:: kernels synthesise inputs and run the consistency checks)\n")
    for k, (n, t, u) in agg.items():
        f.write("%-112s %5d %14.1f %s\n" % (k, n, t, u))
print(len(rows), "launches")
PY
