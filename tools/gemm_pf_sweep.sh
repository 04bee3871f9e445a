#!/bin/bash
# usage (under gpurun): tools/gemm_pf_sweep.sh "0 4 6 10"   -> gpurun_out/gemm_pf_<d>.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for d in $1; do
  UB_GEMM_PF=$d timeout 120 python tools/gemm_check.py > gpurun_out/gemm_pf_$d.log 2>&1
  echo "== PF=$d rc=$?"
  grep "us " gpurun_out/gemm_pf_$d.log | sed 's/max_err.*OK//; s/a_mn=0 b_mn=0 //; s/ split=1://' | cut -c1-150
done
