#!/bin/bash
# A/B of the warp-per-instance kinematics kernels' register budgets: GPU tests + default bench stage split on the default
# library, config 2 (Acrobot, 2^20 instances) on every library given as argument
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stage_ms'], d['e2e']['value'])"
for L in qpcontrol.jl_b200/csrc/libqpcontrol_b200.so "$@"; do
  echo $L; QPC_LIB_PATH=$PWD/$L timeout 300 python tools/bench_configs.py --only-acrobot 2>&1 | tail -1 | cut -c1-260
done
echo "QPC_KIN_WARP=0"; QPC_KIN_WARP=0 timeout 300 python tools/bench_configs.py --only-acrobot 2>&1 | tail -1 | cut -c1-260
