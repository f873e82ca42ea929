#!/bin/bash
set -u
mkdir -p gpurun_out
for T in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $T --print-limit 20 python tools/sanitize_small.py > gpurun_out/san_$T.log 2>&1; echo "$T rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|atlas|dense|acrobot" gpurun_out/san_$T.log | head -12
done
