#!/bin/bash
# ncu visit: launch list + full capture of one (unfused) step
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 30 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-breakdown > gpurun_out/ncu_launch.log 2>&1 ; echo "ncu1 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_(loss_fwd|loss_bwd|analysis|synthesis|mask)' -s 33 -c 11 -o gpurun_out/prof -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-breakdown > gpurun_out/ncu_full.log 2>&1 ; echo "ncu2 rc=$?" ; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
