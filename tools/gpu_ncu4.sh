#!/bin/bash
# ncu full capture of the ADMM kernel limited to ONE CTA per SM (latency chain without co-resident CTAs); args: TAG
set -u
mkdir -p gpurun_out
TAG=${1:-s}
QPC_ADMM_SMEM_PAD=150000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:qpc_admm -s 5 -c 1 -f -o gpurun_out/${TAG}_admm1 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 65536 > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
