#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-e}
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -n 2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","stage_ms","accepted_frac")}, d["roofline"]["iters_mean"], d["roofline"]["factorizations_mean"], d["e2e"]["value"], d["sequential_ticks"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qpc_admm -s 5 -c 1 -f -o gpurun_out/${TAG}_admm \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 65536 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
