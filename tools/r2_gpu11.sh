#!/bin/bash
# round 2, GPU call 11: what the driver runs at round end -- smoke(), the reference arm, the default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
( time python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 ) > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"; cat gpurun_out/r2_bench_reference.json | cut -c1-600; tail -4 gpurun_out/r2_bench_reference.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err; echo "bench rc=$?"; tail -4 gpurun_out/r2_bench_final_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final_n1.json').read().strip().splitlines()[-1])
print("value",d["value"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"], "launches",d["gpu_launches"], d["clocks"])
print("sustained", d["sustained"]["value"], d["sustained"]["roofline"]["frac"], d["sustained"]["clocks"])
for c in d["configs"]:
    if "error" in c: print(c); continue
    print("%-18s %-8s %7.1f GS/s frac %.3f steps %d ok=%s clk=%s %s e2e=%s" % (c["workload"], c["phase"][:8], c["value"], c["roofline"]["frac"], c["steps"], c["parity_spot_check"], c["clocks"]["sm_mhz"], c["clocks"]["reasons"], (c.get("e2e") or {}).get("value")))
PY
