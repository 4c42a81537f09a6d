#!/bin/bash
# Development aid: the ncu evidence of one bench configuration -- launch list of the bench command, then one --set full capture of
# the bench's own closest-point and ray launches.  Usage: tools/ncu_bench.sh <tag>   (writes gpurun_out/<tag>_*)
tag=${1:-r01v}
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_|DeviceRadixSort|DeviceScan' -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_closest -s 1 -c 1 -f -o gpurun_out/${tag}_closest_bench_full $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ray -s 1 -c 1 -f -o gpurun_out/${tag}_ray_bench_full $B > /dev/null 2>&1
ls -la gpurun_out/${tag}_*
