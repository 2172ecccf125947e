# tools/run_sanitize.sh -- under gpurun: compute-sanitizer over small tests that reach every kernel family
mkdir -p gpurun_out
K="ragged or short_late_stages and iw8 or exhaustive_8bit or quadtbl or per_sample_inputs and shipped or full_phase_sweep and cfg0 or nco_chunks"
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$? $(grep -c PASSED gpurun_out/sanitize_$tool.log) $(tail -3 gpurun_out/sanitize_$tool.log | tr '\n' ' ')"
done
