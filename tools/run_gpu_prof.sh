#!/bin/bash
# ncu visit: launch list + full capture of one step of the default composition (stft, mask+istft tail, MR-STFT loss)
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 33 -c 22 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-breakdown > gpurun_out/ncu_launch.log 2>&1 ; echo "ncu1 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_(loss_fwd|loss_bwd|analysis|mask_istft)' -s 27 -c 9 -o gpurun_out/prof -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-breakdown > gpurun_out/ncu_full.log 2>&1 ; echo "ncu2 rc=$?" ; tail -2 gpurun_out/ncu_full.log
timeout 900 ncu --set full --clock-control none -k regex:'k_(analysis|synthesis|mask_fwd|mask_bwd)' -s 15 -c 5 -o gpurun_out/prof_dropin -f python bench.py --fused 0 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-breakdown > gpurun_out/ncu_full2.log 2>&1 ; echo "ncu3 rc=$?" ; tail -2 gpurun_out/ncu_full2.log
ls -la gpurun_out
