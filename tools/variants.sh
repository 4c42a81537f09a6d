#!/bin/bash
# Development aid: runs tools/quick_perf.py <mode> once per prebuilt library variant in build/libtgq_*.so.
mode=${1:-short}
cp thea_b200/libtgq.so /tmp/libtgq_orig.so
for f in build/libtgq_*.so; do
  echo "== $f"; cp $f thea_b200/libtgq.so; timeout 300 python tools/quick_perf.py $mode
done
cp /tmp/libtgq_orig.so thea_b200/libtgq.so
