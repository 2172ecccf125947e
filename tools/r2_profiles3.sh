#!/bin/bash
# round 2: --set full captures of the byte-table kernel with merged records (NCO cfg4 step; random-phase cfg1)
set -u
mkdir -p gpurun_out
cap2() {
  tag=$1; kern=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"$kern" -s $skip -c 1 -f -o gpurun_out/prof_${tag} \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs --no-sustained "$@" > gpurun_out/prof_${tag}.log 2>&1
  ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/prof_${tag}_raw.csv "ncu --set full --clock-control none, python bench.py --steps 2 --warmup 3 $* (launch 4 of the kernel)" > gpurun_out/prof_${tag}.md 2>/dev/null
}
cap2 r2_nco_cfg4_merged 'k_rotate_seededILi8ELi2ELb1ELi4ELi0ELb0E' 3 --workload nco_cfg1
cap2 r2_rotate_cfg1_random_merged 'k_rotate_seededILi8ELi0ELb1ELi4ELi0ELb0E' 3 --workload rotate_cfg1 --phase random
cap2 r2_nco_sintable 'k_lut_smemILb0ELb0ELb0E' 3 --workload nco_sintable_p17
rm -f gpurun_out/*.ncu-rep
for t in r2_nco_cfg4_merged r2_rotate_cfg1_random_merged r2_nco_sintable; do echo "== $t"; grep "Kernel Name\|time_duration\|dram__bytes\|issue_active.avg.pct\|lsu_wavefronts.sum.pct\|wavefronts_mem_shared.sum \|pipe_alu\|fmaheavy" gpurun_out/prof_$t.md; done
