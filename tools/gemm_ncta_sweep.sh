#!/bin/bash
# usage (under gpurun): tools/gemm_ncta_sweep.sh "2 4"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for d in $1; do
  UB_GEMM_NCTA=$d timeout 120 python tools/gemm_check.py > gpurun_out/gemm_ncta_$d.log 2>&1
  echo "== NCTA=$d rc=$?"
  head -1 gpurun_out/gemm_ncta_$d.log
  grep -v "us " gpurun_out/gemm_ncta_$d.log | grep -v "OK$" | head -5
  grep "us " gpurun_out/gemm_ncta_$d.log | sed 's/max_err=\([^ ]*\).*rel=\([^ ]*\) /rel=\2 /; s/a_mn=0 b_mn=0 //; s/ split=1://' | cut -c1-170
done
