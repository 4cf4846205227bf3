#!/bin/bash
# tensor-core DFT-as-GEMM alternative: accuracy + CUDA-event times, then ncu tensor-pipe utilisation of the same launches
set -u
mkdir -p gpurun_out
python tools/tc_dft_bench.py > gpurun_out/tc_dft.json 2> gpurun_out/tc_dft.err; echo rc=$?; cat gpurun_out/tc_dft.json; tail -3 gpurun_out/tc_dft.err
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:'k_tc_stft512|k_analysis' -c 6 --csv --log-file gpurun_out/tc_dft_ncu.csv python tools/tc_dft_bench.py > /dev/null 2>&1; echo ncu rc=$?; cat gpurun_out/tc_dft_ncu.csv | tail -60
