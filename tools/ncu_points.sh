#!/bin/bash
# Development aid: cfg2-shaped k-NN / sample-graph / filtered 24-NN timings with candidate buffers (1) and sorted sets (0), then the
# ncu counters of the k-NN and graph launches of both.  Usage: tools/ncu_points.sh <tag>
tag=${1:-r02r}
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__registers_per_thread
for b in 1 0; do
  TGQ_SET_BUFFER=$b python tools/prof_points.py > gpurun_out/${tag}_points_buf$b.log 2>&1
  TGQ_SET_BUFFER=$b timeout 600 ncu --metrics $M --clock-control none -k 'regex:k_knn|k_ball_graph' -c 6 --csv --log-file gpurun_out/${tag}_points_ncu_buf$b.csv python tools/prof_points.py > /dev/null 2>&1
done
cat gpurun_out/${tag}_points_buf*.log
