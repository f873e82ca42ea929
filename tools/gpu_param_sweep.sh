#!/bin/bash
# ADMM-kernel time and iteration statistics of the one-warp kernel over its schedule parameters (env knobs of api.cu: warp_params)
for S in notebook test_suite; do
for cfg in "30 1.35 25 25 25" "30 1.35 20 25 25" "30 1.35 30 25 25" "20 1.35 25 25 25" "40 1.35 25 25 25" "30 1.5 25 25 25" "30 1.25 25 25 25" "30 1.35 25 25 50" "30 1.35 50 25 25" "35 1.35 25 25 25"; do
  set -- $cfg
  echo -n "$S kappa=$1 growth=$2 first=$3 check=$4 aitken=$5 : "
  QPC_WARP_KAPPA=$1 QPC_WARP_GROWTH=$2 QPC_WARP_FIRST=$3 QPC_WARP_CHECK=$4 QPC_WARP_AITKEN=$5 python tools/one_tick.py $S 16384 4 2>/dev/null | grep stage | sed 's/.*stage ms asm.admm.id = //'
done; done
