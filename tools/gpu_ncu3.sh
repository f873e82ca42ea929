#!/bin/bash
# ncu full captures of the assembly and inverse-dynamics kernels (one 16384-instance launch each); args: TAG
set -u
mkdir -p gpurun_out
TAG=${1:-k}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qpc_assemble -s 5 -c 1 -f -o gpurun_out/${TAG}_asm \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 65536 > gpurun_out/${TAG}_ncu_asm.log 2>&1; echo "ncu asm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qpc_inverse -s 5 -c 1 -f -o gpurun_out/${TAG}_id \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 65536 > gpurun_out/${TAG}_ncu_id.log 2>&1; echo "ncu id rc=$?"
