#!/bin/bash
# A/B bench of library variants: args = TAG lib1 lib2 ...
set -u
mkdir -p gpurun_out
TAG=$1; shift
for L in "$@"; do
  N=$(basename $L .so)
  QPC_LIB_PATH=$PWD/$L timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_${N}.json 2> gpurun_out/${TAG}_${N}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${N}.json").read().strip().splitlines()[-1])
    print("$N", {k:d[k] for k in ("value","ms_per_step","stage_ms","accepted_frac","iters_max")}, d["roofline"]["iters_mean"], d["roofline"]["factorizations_mean"], d["roofline"]["frac"], d["e2e"]["value"])
except Exception as e: print("$N no bench", e)
PY
done
