#!/bin/bash
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 700 compute-sanitizer --tool $tool --launch-timeout 0 python tests/sanitizer_smoke.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer smoke done" gpurun_out/sanitizer_$tool.log | tail -3
done
# the opt-in saved-spectrum loss backward
SE_MRSTFT_SAVE_SPECTRUM=1 timeout 600 compute-sanitizer --tool memcheck --launch-timeout 0 python tests/sanitizer_smoke.py > gpurun_out/sanitizer_memcheck_saved.log 2>&1
echo "memcheck (saved spectrum) rc=$?"; grep -E "ERROR SUMMARY|sanitizer smoke done" gpurun_out/sanitizer_memcheck_saved.log | tail -2
# the alternative engines (signal-pair packed fp32, two-pass): same smoke, memcheck + racecheck
for eng in 2 3; do for tool in memcheck racecheck; do
  SE_ENGINE=$eng timeout 700 compute-sanitizer --tool $tool --launch-timeout 0 python tests/sanitizer_smoke.py > gpurun_out/sanitizer_${tool}_eng$eng.log 2>&1
  echo "$tool engine $eng rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer smoke done" gpurun_out/sanitizer_${tool}_eng$eng.log | tail -3
done; done
