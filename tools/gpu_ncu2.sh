#!/bin/bash
# GPU tests + ncu full capture of one 16384-instance ADMM launch (batch 65536 = 4 chunks of 16384); args: TAG
set -u
mkdir -p gpurun_out
TAG=${1:-p}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qpc_admm -s 5 -c 1 -f -o gpurun_out/${TAG}_admm \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 65536 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/${TAG}_ncu_full.log
