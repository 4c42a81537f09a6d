"""Key counters of an `ncu --page raw --csv` export for the traversal kernels (development aid)."""
import csv, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg',
        'l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed', 'l1tex__f_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'gpu__compute_memory_access_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.avg']
def main(path):
    rows = list(csv.reader(open(path))); hdr, units, data = rows[0], rows[1], rows[2:]
    for r in data:
        print(r[hdr.index('Kernel Name')][:70])
        for k in KEYS:
            if k in hdr: print(f"  {k:90s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
        for i, h in enumerate(hdr):
            if 'wavefront' in h and h not in KEYS and r[i] not in ('', 'n/a'): print(f"  {h:90s} {r[i]} {units[i]}")
if __name__ == '__main__':
    main(sys.argv[1])
