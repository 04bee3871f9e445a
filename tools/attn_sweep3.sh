#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/attn_sweep3.log
: > $L
for shp in "2 17 2" "3 128 4" "2 160 3" "4 197 12" "2 256 3" "3 288 5" "2 300 3" "3 320 12" "32 320 12 time" "32 320 16 time"; do
  echo "=== $shp" >> $L
  timeout 90 python tools/attn_bwd_check.py $shp >> $L 2>&1
  echo "rc=$?" >> $L
done
echo "=== adversarial" >> $L
timeout 90 python tools/attn_fwd_adv.py >> $L 2>&1
echo "rc=$?" >> $L
echo "=== mma.sync forward for comparison" >> $L
UB_ATTN_FWD_LSE_TC=0 timeout 90 python tools/attn_bwd_check.py 32 320 12 time >> $L 2>&1
grep -v " OK$" $L
