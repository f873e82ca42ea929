#!/bin/bash
# ncu full capture of the ADMM kernel (one launch) + launch list; args: TAG [BATCH]
set -u
mkdir -p gpurun_out
TAG=${1:-p}; BATCH=${2:-4096}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qpc_admm -s 3 -c 1 -f -o gpurun_out/${TAG}_admm \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch $BATCH > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
