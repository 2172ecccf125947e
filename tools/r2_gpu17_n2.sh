#!/bin/bash
# round 2, GPU call 17 (2 GPUs): the final bench command under torchrun, as the driver's SCALE run launches it
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r2_final_bench_n2.json 2> gpurun_out/r2_final_bench_n2.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_final_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench_n2.json').read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"].get("value"), d["e2e"].get("error"), "xchg",{k:v for k,v in d["scatter_gather"].items() if k in ("value","transport","parity","error")})
print("sustained", d["sustained"]["value"])
for c in d["configs"]: print(c.get("workload"), c.get("phase","")[:8], c.get("value"), c.get("parity_spot_check"), c.get("error"))
PY
( time python bench.py --impl reference --gpus 2 --steps 1 --warmup 1 ) 2>&1 | cut -c1-200 | tail -5
