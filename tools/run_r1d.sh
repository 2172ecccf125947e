# tools/run_r1d.sh -- under gpurun: GPU tests, sanitizer pass over small tests, ncu evidence for the changed kernels
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r1d_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1d_pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "ragged or short_late_stages and iw8 or exhaustive_8bit or smoke" > gpurun_out/r1d_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -5 gpurun_out/r1d_sanitizer.log
bash tools/ncu_capture.sh topolar_tail --workload topolar_cfg2 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/prof_topolar_tail_raw.csv "k_topolar<21,10> cfg2 (short late stages)" > gpurun_out/r1d_ncu_topolar_tail.md
bash tools/ncu_final.sh > gpurun_out/r1d_ncu_final.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_final_raw.csv "k_rotate_seeded<8,0,1,0> cfg1 sweep (words), after the r1d trims" > gpurun_out/r1d_ncu_seeded.md
cat gpurun_out/r1d_ncu_topolar_tail.md gpurun_out/r1d_ncu_seeded.md | grep -v "^$" | cut -c1-150
rm -f gpurun_out/*.ncu-rep
