#!/bin/bash
# round 2, GPU call 5 (2 GPUs): multi-device tests, copy probe, NCCL/peer scatter-gather, torchrun bench at N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_n2_topo.txt 2>&1; nproc >> gpurun_out/r2_n2_topo.txt; lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/r2_n2_topo.txt; free -g >> gpurun_out/r2_n2_topo.txt
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_sharding.py -q > gpurun_out/r2_n2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_n2_pytest.log
tail -15 gpurun_out/r2_n2_pytest.log
for g in 1 2; do timeout 300 ./cordic_b200/zcordic_bench --pcie-probe -g $g -l 28 --json; done > gpurun_out/r2_n2_pcie_probe.txt 2>&1; cat gpurun_out/r2_n2_pcie_probe.txt
timeout 600 ./cordic_b200/zcordic_bench -g 2 --scatter -l 28 -s 5 --json > gpurun_out/r2_n2_scatter.txt 2>&1; cat gpurun_out/r2_n2_scatter.txt
for c in 4 16 32; do timeout 600 ./cordic_b200/zcordic_bench -g 2 --scatter -l 28 -s 5 --chunks $c --transport nccl --json; done > gpurun_out/r2_n2_scatter_chunks.txt 2>&1; cat gpurun_out/r2_n2_scatter_chunks.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_n2_bench.json 2> gpurun_out/r2_n2_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_n2_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_n2_bench.json').read().strip().splitlines()[-1])
    print("value",d["value"],"e2e",d["e2e"],"xchg",d["scatter_gather"])
    for c in d["configs"]: print(c.get("workload"), c.get("phase"), c.get("value"), c.get("parity_spot_check"), c.get("error"))
except Exception as e: print("parse failed", e)
PY
