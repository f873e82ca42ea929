#!/bin/bash
# final pass: tests, smoke, default bench, masks bench, reference arm
set -u
mkdir -p gpurun_out
TAG=${1:-z}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --masks --steps 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_masks.json 2> gpurun_out/${TAG}_bench_masks.err; echo "bench masks rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
for f in ("bench","bench_masks","bench_ref"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_%s.json"%f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","accepted_frac","gpu_launches","latency_us_by_batch")}, d.get("e2e",{}).get("value"))
    except Exception as e: print(f,"failed",e)
PY
