#!/bin/bash
# ncu --set full of every transform-type kernel, pair engine and scalar engine (second repetition: warm tables).
# Reports are converted to CSV on the box (gpurun_out/ is capped at 64 MiB); only the pair-engine report travels.
set -u
mkdir -p gpurun_out
for eng in 2 1; do
  SE_ENGINE=$eng timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(analysis|synthesis|loss_fwd|loss_bwd)' -s 10 -c 10 -o /tmp/prof_eng$eng -f python tools/prof_ops.py > gpurun_out/ncu_eng$eng.log 2>&1 ; echo "ncu eng$eng rc=$?" ; tail -2 gpurun_out/ncu_eng$eng.log
  ncu -i /tmp/prof_eng$eng.ncu-rep --page raw --csv > gpurun_out/raw_eng$eng.csv 2>/dev/null
done
cp /tmp/prof_eng2.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out
