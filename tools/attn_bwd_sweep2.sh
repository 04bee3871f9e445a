#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/attn_bwd_sweep.log
: > $L
for shp in "2 17 2" "3 288 5" "2 300 3" "32 320 12 time" "32 320 16 time"; do
  echo "=== $shp" >> $L
  timeout 90 python tools/attn_bwd_check.py $shp >> $L 2>&1
  echo "rc=$?" >> $L
done
grep -v "OK$" $L
