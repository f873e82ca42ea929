#!/bin/bash
# round-end pass: tests, smoke, bench (+ reference arm, masks), config sweep, ncu launch list of the bench command; args: TAG
set -u
mkdir -p gpurun_out
TAG=${1:-end}
bash tools/gpu_full.sh $TAG
timeout 900 python bench.py --masks --steps 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_masks.json 2> gpurun_out/${TAG}_bench_masks.err; echo "bench masks rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
python - <<PY
import json
for f in ("bench","bench_masks","bench_ref"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_%s.json"%f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","accepted_frac","gpu_launches","stage_ms")}, d.get("e2e",{}).get("value"), (d.get("roofline") or {}).get("frac"))
    except Exception as e: print(f,"failed",e)
PY
