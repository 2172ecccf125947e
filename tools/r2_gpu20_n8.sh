#!/bin/bash
# round 2, GPU call 20 (8 GPUs): the final bench command under torchrun at N = 8, as the driver's SCALE run launches it
mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r2_final_bench_n8.json 2> gpurun_out/r2_final_bench_n8.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_final_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench_n8.json').read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"].get("value"), d["e2e"].get("error"), "xchg",{k:(v if not isinstance(v,dict) else v.get("value")) for k,v in d["scatter_gather"].items() if k in ("value","transport","parity","error","nccl","peer","copy")})
print("sustained", d["sustained"]["value"])
for c in d["configs"]: print(c.get("workload"), c.get("phase","")[:8], c.get("value"), c.get("parity_spot_check"), c.get("error"))
PY
