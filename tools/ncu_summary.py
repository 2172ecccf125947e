#!/usr/bin/env python3
"""tools/ncu_summary.py <raw.csv> [title] -- prints the handful of ncu metrics the design cares about as a
markdown table (input: `ncu -i X.ncu-rep --page raw --csv`)."""
import csv
import sys

KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.per_cycle_active', 'sm__inst_executed.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_requests_srcunit_tex_op_write.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'sm__cycles_elapsed.avg']


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    if len(sys.argv) > 2:
        print("# " + sys.argv[2] + "\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in m:
            print("| %s | %s | %s |" % (k, m[k][0], m[k][1]))


if __name__ == "__main__":
    main()
