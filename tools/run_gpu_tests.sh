#!/bin/bash
# usage: tools/run_gpu_tests.sh [pytest args]   (runs under gpurun on the B200 box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -s "$@" > gpurun_out/pytest_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
