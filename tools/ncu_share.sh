#!/bin/bash
# Development aid: ONE rank's share of the cfg5 batch at world size 8, on one GPU (tools/shard_probe.py) -- the launch list of its two
# C-ABI calls (selection, sort, traversal) and one --set full capture of the share's closest-point launch.  Multi-rank commands cannot
# run under ncu; this is the per-rank evidence for the N = 8 numbers.   Usage: tools/ncu_share.sh <tag>
tag=${1:-r02y}
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_|DeviceRadixSort|DeviceScan|DeviceSelect' -c 600 --csv --log-file gpurun_out/${tag}_share8_launches.csv python tools/shard_probe.py 8 > gpurun_out/${tag}_share8_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_closest_pool -s 2 -c 1 -f -o gpurun_out/${tag}_share8_closest_full python tools/shard_probe.py 8 > /dev/null 2>&1
python tools/shard_probe.py 8 > gpurun_out/${tag}_share8_probe.json 2>&1
ls -la gpurun_out/${tag}_share8*
