"""Summarise an `ncu --page raw --csv` export: the handful of counters the design decisions cite."""
import csv, sys, json
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__warps_eligible.avg.per_cycle_active', 'sm__inst_executed.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'local_load bytes', 'smsp__inst_executed_op_local_ld.sum',
        'l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum']
def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for r in data:
        d = {'kernel': r[hdr.index('Kernel Name')][:60]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w); d[w] = f"{r[i]} {units[i]}".strip()
        out.append(d)
    print(json.dumps(out, indent=1))
if __name__ == '__main__':
    main(sys.argv[1])
