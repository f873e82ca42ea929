#!/bin/bash
# refresh of the secondary measurements after a kernel change: masks bench (config 4) and configs 1/2/4/5
set -u
mkdir -p gpurun_out
TAG=${1:-rf}
timeout 900 python bench.py --masks --steps 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_masks.json 2> gpurun_out/${TAG}_bench_masks.err; echo "bench masks rc=$?"
timeout 900 python tools/bench_configs.py --out gpurun_out/${TAG}_configs.json > gpurun_out/${TAG}_configs.log 2>&1; echo "configs rc=$?"
tail -12 gpurun_out/${TAG}_configs.log | cut -c1-330
