#!/bin/bash
# Development aid: a few ncu counters of the closest kernel for every prebuilt library variant in build/.
cp thea_b200/libtgq.so /tmp/libtgq_orig.so
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
for f in build/libtgq_*.so; do
  echo "== $f"; cp $f thea_b200/libtgq.so
  timeout 200 ncu --metrics $M --clock-control none -k regex:k_closest -s 2 -c 1 --csv python tools/prof_one.py closest 2236 4000000 2>/dev/null | grep -E "k_closest" | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"'
done
cp /tmp/libtgq_orig.so thea_b200/libtgq.so
