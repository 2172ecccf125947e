#!/bin/bash
# round 2 profile collection (1 GPU, under gpurun): launch list of the default bench command, --set full captures of
# the kernels round 2 added or changed, compute-sanitizer over the new paths.  Summaries land in gpurun_out/.
set -u
mkdir -p gpurun_out
# 1. launch list: the default command's timed regions (headline + every extra configuration at its minimum step count)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_default.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --sustained-steps 5 --configs-seconds 0 > gpurun_out/r2_launches_default.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2_launches_default.csv')) if len(r) > 5 and r[0].strip('"').isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4]; val = float(r[-1].replace(',', ''))
    unit = r[-2]
    k = name[:110]
    a = agg.setdefault(k, [0, 0.0, unit]); a[0] += 1; a[1] += val
with open('gpurun_out/r2_launches_default.txt', 'w') as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --sustained-steps 5 --configs-seconds 0\n")
    f.write("# kernel | launches | total duration | unit (cold-cache, serialised)\n")
    for k, (n, t, u) in agg.items():
        f.write("%-112s %5d %14.1f %s\n" % (k, n, t, u))
PY
cat gpurun_out/r2_launches_default.txt | head -60
# 2. --set full captures
cap() { tag=$1; kern=$2; skip=$3; shift 3; NCU_KERNEL="$kern" NCU_SKIP=$skip bash tools/ncu_capture.sh $tag "$@" > /dev/null 2>&1; }
cap2() { # capture only (no launch list)
  tag=$1; kern=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$kern" -s $skip -c 1 -f -o gpurun_out/prof_${tag} \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs --no-sustained "$@" > gpurun_out/prof_${tag}.log 2>&1
  ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/prof_${tag}_raw.csv "ncu --set full, bench.py $* (kernel regex $kern)" > gpurun_out/prof_${tag}.md 2>/dev/null
}
cap2 r2_rotate_cfg1 'k_rotate_seeded<8, 0, 1, 3, 0, 0>' 3 --workload rotate_cfg1
cap2 r2_rotxy_words 'k_rotate_dirs<8, 1, 1, 1>' 3 --workload rotate_xy_cfg1
cap2 r2_rotxy_random 'k_rotate_dirs<8, 1, 1, 0>' 3 --workload rotate_xy_cfg1 --phase random --seed-mode packed
cap2 r2_nco_comb 'k_rotate_seeded<8, 2, 1, 3, 1, 0>' 3 --workload nco_cfg1 --nco-step 0x80000001
cap2 r2_nco_cfg4_packed 'k_rotate_seeded<8, 2, 1, 2, 0, 0>' 3 --workload nco_cfg1
cap2 r2_rotate_o16 'k_rotate_seeded<3, 0, 1, 3, 0, 1>' 3 --workload rotate_o16_cfg0
cap2 r2_topolar_i16 'k_topolar<21, 10, 1>' 3 --workload topolar_i16_cfg2
rm -f gpurun_out/*.ncu-rep
for t in r2_rotate_cfg1 r2_rotxy_words r2_rotxy_random r2_nco_comb r2_nco_cfg4_packed r2_rotate_o16 r2_topolar_i16; do echo "== $t"; grep "Kernel Name\|time_duration\|dram__bytes\|issue_active.avg.pct\|lsu_wavefronts.sum.pct\|wavefronts_mem_shared.sum \|pipe_alu\|fmaheavy" gpurun_out/prof_$t.md; done
# 3. compute-sanitizer over the round-2 paths (small sizes via the tests that use them)
K="comb or i16 or o16 or host_multi or word_suffix and shipped"
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_round2.py tests/test_gpu_runtime.py -x -q -k "$K" > gpurun_out/r2_sanitize_$tool.log 2>&1
  echo "$tool rc=$? $(tail -3 gpurun_out/r2_sanitize_$tool.log | tr '\n' ' ')"
done
