#!/bin/bash
# round 2: the whole gpu tier under other seeds of the randomised tests (random configurations through every kernel flavour)
mkdir -p gpurun_out
for seed in 11 12 13 14; do echo -n "ZC_TEST_SEED=$seed: "; ZC_TEST_SEED=$seed timeout 900 python -m pytest tests -m gpu -q -x -k "not full_size and not rtl" 2>&1 | tail -1; done > gpurun_out/r2_soak_full.txt
cat gpurun_out/r2_soak_full.txt
