#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-c}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python tools/bench_configs.py --out gpurun_out/${TAG}_configs.json > gpurun_out/${TAG}_configs.log 2>&1; echo "configs rc=$?"
tail -12 gpurun_out/${TAG}_configs.log | cut -c1-330
