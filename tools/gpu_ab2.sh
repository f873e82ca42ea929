#!/bin/bash
# parity tests on the default build, then A/B bench of library variants: args = TAG lib1 lib2 ...
set -u
mkdir -p gpurun_out
TAG=$1; shift
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${TAG}_pytest_gpu.log
bash tools/gpu_ab.sh $TAG "$@"
