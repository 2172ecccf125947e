#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -q -k "scatter or multi" 2>&1 | tail -2
timeout 300 ./cordic_b200/zcordic_bench -g 2 --scatter -l 26 -s 3 --json 2>&1 | grep "^{" | cut -c1-330
