#!/bin/bash
# Round-2 evidence run on one B200: tests, both bench lines + the reference arm, launch list, ncu captures, configs.
set -u
mkdir -p gpurun_out
T=r2z
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -n 2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_notebook.json 2> gpurun_out/${T}_bench_notebook.err
timeout 600 python bench.py --steps 20 --warmup 3 --settings test_suite > gpurun_out/${T}_bench_test_suite.json 2> gpurun_out/${T}_bench_test_suite.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qpc_admm_warp -s 1 -c 1 -f -o gpurun_out/${T}_warp python tools/one_tick.py notebook 16384 2 > gpurun_out/${T}_warp_ncu.log 2>&1; echo "ncu warp rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qpc_tiny -s 1 -c 1 -f -o gpurun_out/${T}_tiny python tools/acrobot_tick.py 1048576 2 > gpurun_out/${T}_tiny_ncu.log 2>&1; echo "ncu tiny rc=$?"
timeout 900 python tools/bench_configs.py --out gpurun_out/${T}_configs.json > gpurun_out/${T}_configs.log 2>&1; echo "configs rc=$?"
python - <<PY
import json
for n in ("notebook", "test_suite"):
    d = json.loads(open("gpurun_out/${T}_bench_%s.json" % n).read().strip().splitlines()[-1])
    print(n, d["value"], d["ms_per_step"], d["stage_ms"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "cpu", d.get("cpu_baseline", {}).get("value"))
PY
