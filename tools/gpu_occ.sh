#!/bin/bash
# ADMM kernel time vs CTAs per SM (dynamic shared memory padded to 1 / 2 / 3 CTAs per SM)
set -u
mkdir -p gpurun_out
for PAD in 0 40000 150000; do
  QPC_ADMM_SMEM_PAD=$PAD timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/occ_$PAD.json 2> gpurun_out/occ_$PAD.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/occ_$PAD.json").read().strip().splitlines()[-1])
    print("pad $PAD", d["stage_ms"], d["ms_per_step"])
except Exception as e: print("pad $PAD failed", e)
PY
done
