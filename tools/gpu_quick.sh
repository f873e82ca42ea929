#!/bin/bash
# quick GPU iteration: parity tests (all, no -x) and a short bench
set -u
mkdir -p gpurun_out
TAG=${1:-q}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","stage_ms","accepted_frac","iters_max")}, d["roofline"]["iters_mean"], d["roofline"]["factorizations_mean"], d["roofline"]["frac"], d["e2e"]["value"])
except Exception as e: print("no bench", e)
PY
