M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers
for p in 0 1; do
  TGQ_POOL=$p timeout 300 ncu --metrics $M --clock-control none -k regex:k_closest -s 2 -c 1 --csv --log-file gpurun_out/r01t_ncu_pool$p.csv python tools/prof_one.py closest 2236 8000000 > gpurun_out/r01t_prof_pool$p.log 2>&1
done
python - <<'PY'
import csv
for p in (0,1):
    rows=[r for r in csv.reader(open(f"gpurun_out/r01t_ncu_pool{p}.csv")) if len(r)>5]
    hdr=rows[0]
    print("pool",p)
    for r in rows[1:]:
        print("  ", r[hdr.index("Metric Name")], r[hdr.index("Metric Value")], r[hdr.index("Metric Unit")])
PY
