#!/bin/bash
# round-2 evidence: ncu launch list of the bench command + --set full of one step's kernels (CSV exported on the box)
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 20 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --rounds 1 --no-e2e --no-cpu-baseline --no-breakdown --no-configs --no-incumbent > gpurun_out/ncu_launch.log 2>&1 ; echo "ncu1 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_(loss_fwd|loss_bwd|analysis|mask_istft)' -s 27 -c 9 -o /tmp/prof_r02 -f python bench.py --steps 1 --warmup 3 --rounds 1 --no-e2e --no-cpu-baseline --no-breakdown --no-configs --no-incumbent > gpurun_out/ncu_full.log 2>&1 ; echo "ncu2 rc=$?" ; tail -2 gpurun_out/ncu_full.log
ncu -i /tmp/prof_r02.ncu-rep --page raw --csv > gpurun_out/r02_raw_step.csv 2>/dev/null
ls -la gpurun_out | tail -8
