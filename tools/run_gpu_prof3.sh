#!/bin/bash
# source-level stall sampling of the default engine: k_analysis (stft) and the loss kernels at n = 1024
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(analysis|loss_fwd|loss_bwd)' -s 8 -c 7 -o gpurun_out/prof_src -f python tools/prof_ops.py > gpurun_out/ncu_src.log 2>&1 ; echo "rc=$?" ; tail -2 gpurun_out/ncu_src.log; ls -la gpurun_out/prof_src.ncu-rep
