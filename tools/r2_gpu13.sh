#!/bin/bash
# round 2, GPU call 13: full gpu tier with the new tests, three seeds of the randomised round-2 tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest13.log
tail -5 gpurun_out/r2_pytest13.log
for seed in 1 2 3; do ZC_TEST_SEED=$seed timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "not full_size" 2>&1 | tail -1; done > gpurun_out/r2_soak.txt
cat gpurun_out/r2_soak.txt
