#!/bin/bash
# round 2, GPU call 7 (8 GPUs): multi-device tests, copy probe at 4/8, scatter-gather at 4/8, torchrun bench at N=8
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_n8_topo.txt 2>&1; nproc >> gpurun_out/r2_n8_topo.txt; lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/r2_n8_topo.txt; free -g >> gpurun_out/r2_n8_topo.txt
timeout 600 python -m pytest tests/test_gpu_round2.py -q -k "multi or scatter" > gpurun_out/r2_n8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_n8_pytest.log
tail -5 gpurun_out/r2_n8_pytest.log
for g in 1 2 4 8; do timeout 300 ./cordic_b200/zcordic_bench --pcie-probe -g $g -l 28 --json; done > gpurun_out/r2_n8_pcie_probe.txt 2>&1; cat gpurun_out/r2_n8_pcie_probe.txt
for g in 2 4 8; do timeout 600 ./cordic_b200/zcordic_bench -g $g --scatter -l 28 -s 5 --json 2>&1 | grep "^{"; done > gpurun_out/r2_n8_scatter.txt; cat gpurun_out/r2_n8_scatter.txt
timeout 600 ./cordic_b200/zcordic_bench -g 8 --scatter -l 28 -s 5 --chunks 4 --transport nccl --json 2>&1 | grep "^{" >> gpurun_out/r2_n8_scatter.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_n8_bench.json 2> gpurun_out/r2_n8_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_n8_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_n8_bench.json').read().strip().splitlines()[-1])
    print("value",d["value"],"e2e",d["e2e"],"xchg",d["scatter_gather"])
    for c in d["configs"]: print(c.get("workload"), c.get("phase"), c.get("value"), c.get("parity_spot_check"), c.get("error"))
except Exception as e: print("parse failed", e)
PY
