#!/bin/bash
# round 2, GPU call 12 (8 GPUs): the three scatter/gather transports at N = 8 and N = 2 after the mutual peer mapping
mkdir -p gpurun_out
for g in 8 2; do timeout 600 ./cordic_b200/zcordic_bench -g $g --scatter -l 28 -s 5 --json 2>&1 | grep "^{"; done > gpurun_out/r2_n8_scatter3.txt
timeout 600 ./cordic_b200/zcordic_bench -g 8 --scatter -l 28 -s 5 --transport copy --json 2>&1 | grep "^{" >> gpurun_out/r2_n8_scatter3.txt
python - <<'PY'
import json
for l in open('gpurun_out/r2_n8_scatter3.txt'):
    d=json.loads(l); print(d["n_gpus"], d["transport"], {k:(d[k]["value"], d[k]["dev0_ingress_gbs"], d[k]["parity"]) for k in ("nccl","peer","copy") if k in d})
PY
