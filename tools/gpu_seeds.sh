#!/bin/bash
set -u
mkdir -p gpurun_out
for R in 0 1 2 3 4 5 6 7; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --as-rank $R > gpurun_out/seed_$R.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/seed_$R.json").read().strip().splitlines()[-1])
print("rank-seed $R", round(d["ms_per_step"],3), d["stage_ms"]["admm"], d["roofline"]["iters_mean"], d["iters_max"], d["accepted_frac"])
PY
done
