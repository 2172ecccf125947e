#!/bin/bash
# round 2, GPU call 8 (4 GPUs): scatter/gather transports and NCCL channel settings
mkdir -p gpurun_out
x() { timeout 300 ./cordic_b200/zcordic_bench -g 4 --scatter -l 28 -s 5 --json "$@" 2>&1 | grep "^{" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1:], {k:(d[k]["value"], d[k]["dev0_ingress_gbs"], d[k]["parity"]) for k in ("nccl","peer","copy") if k in d})' "$@"; }
{
x --transport both
x --transport copy --chunks 4
x --transport copy --chunks 16
x --transport nccl --chunks 16
NCCL_MIN_P2P_NCHANNELS=8 x --transport nccl
NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32 x --transport nccl
NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32 x --transport nccl --chunks 16
} > gpurun_out/r2_n4_scatter_tuning.txt 2>&1
cat gpurun_out/r2_n4_scatter_tuning.txt
