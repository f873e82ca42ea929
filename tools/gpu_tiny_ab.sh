#!/bin/bash
# A/B of library variants on the tiny-QP workloads (config 2 and two small dense sizes), then GPU tests + default bench
for L in "$@" qpcontrol.jl_b200/csrc/libqpcontrol_b200.so; do
  echo $L
  QPC_LIB_PATH=$PWD/$L timeout 300 python tools/bench_configs.py --only-acrobot 2>&1 | tail -1 | cut -c1-230
  QPC_LIB_PATH=$PWD/$L timeout 300 python tools/bench_configs.py --only-dense --dense 15,15,65536 2>&1 | tail -1 | cut -c1-150
  QPC_LIB_PATH=$PWD/$L timeout 300 python tools/bench_configs.py --only-dense --dense 6,8,262144 2>&1 | tail -1 | cut -c1-150
done
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stage_ms'], d['e2e']['value'])"
