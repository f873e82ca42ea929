#!/bin/bash
# full GPU pass: tests, smoke, bench (+reference arm), config sweep; args: TAG
set -u
mkdir -p gpurun_out
TAG=${1:-f}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/${TAG}_bench.err
timeout 900 python tools/bench_configs.py --out gpurun_out/${TAG}_configs.json > gpurun_out/${TAG}_configs.log 2>&1; echo "configs rc=$?"
tail -12 gpurun_out/${TAG}_configs.log | cut -c1-400
