#!/bin/bash
for S in notebook test_suite; do
for cfg in "30 1.35 25 5 25" "30 1.35 25 10 25" "30 1.35 25 3 25" "30 1.35 25 25 25"; do
  set -- $cfg
  echo -n "$S check=$4 : "
  QPC_WARP_KAPPA=$1 QPC_WARP_GROWTH=$2 QPC_WARP_FIRST=$3 QPC_WARP_CHECK=$4 QPC_WARP_AITKEN=$5 python tools/one_tick.py $S 16384 4 2>/dev/null | grep stage | sed 's/.*stage ms asm.admm.id = //'
done; done
