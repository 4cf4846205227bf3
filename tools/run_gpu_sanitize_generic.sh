#!/bin/bash
# compute-sanitizer over the smoke script (which now includes the general-geometry path): memcheck, racecheck, initcheck
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 800 compute-sanitizer --tool $tool --launch-timeout 0 python tests/sanitizer_smoke.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer smoke done" gpurun_out/sanitizer_$tool.log | tail -3
done
