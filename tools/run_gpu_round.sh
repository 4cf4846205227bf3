#!/bin/bash
# One GPU-box visit: smoke, gpu tests, bench, ncu launch list + full capture.  Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -8 gpurun_out/smoke.log
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1 ; echo "pytest rc=$?" ; tail -25 gpurun_out/pytest.log
echo "== bench" ; timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "bench rc=$?" ; cat gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ; cat gpurun_out/bench_ref.json


ls -la gpurun_out
echo "== ncu full (unfused drop-in ops)" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(analysis|synthesis|mask)' -s 15 -c 5 -o gpurun_out/prof_unfused -f python bench.py --fused 0 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-breakdown > gpurun_out/ncu_full2.log 2>&1 ; echo "ncu3 rc=$?" ; tail -2 gpurun_out/ncu_full2.log
