#!/bin/bash
# Development aid: key ncu counters of the closest-point kernels (4-wide groups with several leaf-iteration thresholds, and the
# binary pooled kernel) on 8 M queries of the 10 M-triangle mesh, plus un-profiled timings.  Usage: tools/ncu_wide.sh <tag>
tag=${1:-r02e}
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__registers_per_thread
run() {  # name lib wide
  TGQ_LIB=$2 TGQ_WIDE=$3 timeout 300 ncu --metrics $M --clock-control none -k regex:k_closest -s 2 -c 1 --csv --log-file gpurun_out/${tag}_ncu_$1.csv python tools/prof_one.py closest 2236 8000000 > gpurun_out/${tag}_prof_$1.log 2>&1
  TGQ_LIB=$2 TGQ_WIDE=$3 python tools/prof_one.py closest 2236 8000000 >> gpurun_out/${tag}_timings.log 2>&1
}
: > gpurun_out/${tag}_timings.log
for v in "$@"; do :; done
run wide32 "" 1
[ -f build/libtgq_w16.so ] && run wide16 build/libtgq_w16.so 1
[ -f build/libtgq_w8.so ] && run wide8 build/libtgq_w8.so 1
run pool "" 0
for f in build/libtgq_x*.so; do [ -f "$f" ] && run $(basename $f .so | sed s/libtgq_//) $f 1; done
python - <<PY
import csv, glob
for f in sorted(glob.glob("gpurun_out/${tag}_ncu_*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>5]
    if not rows: continue
    hdr=rows[0]
    print(f)
    for r in rows[1:]:
        print("  ", r[hdr.index("Metric Name")], r[hdr.index("Metric Value")], r[hdr.index("Metric Unit")])
PY
cat gpurun_out/${tag}_timings.log
