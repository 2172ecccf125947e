# tools/run_ncu_more.sh -- under gpurun: ncu --set full of the kernels added late in round 1
mkdir -p gpurun_out
cap() { # tag regex bench-args...
  tag=$1; re=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:"$re" -s 2 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs --no-sustained "$@" > gpurun_out/prof_$tag.log 2>&1
  ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/prof_${tag}_raw.csv "$tag: $*" > gpurun_out/ncu_$tag.md
  rm -f gpurun_out/prof_$tag.ncu-rep
}
cap lut_smem_sintable_random 'k_lut_smem' --workload sintable_p17 --phase random
cap lut_smem_quarterwav_random 'k_lut_smem' --workload quarterwav_p18 --phase random
cap quadtbl_rows 'k_quadtbl' --workload quadtbl_p18
cap seedpacked_random 'k_rotate_seeded<8, 0, 1, 2>' --phase random --seed-mode packed
grep -h "Kernel Name\|gpu__time_duration\|dram__bytes\|issue_active.avg.pct\|pipe_alu\|fmaheavy\|wavefronts_mem_shared.sum.pct\|dram_throughput" gpurun_out/ncu_*.md | cut -c1-150
