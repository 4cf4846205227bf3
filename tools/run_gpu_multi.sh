#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err ; echo "bench N=$N rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print("N", d['n_gpus'], "value", round(d['value']), "ms/step", round(d['ms_per_step'],4), "e2e", round(d['e2e']['value']), "loss", d['loss'])
PY
tail -5 gpurun_out/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 | tail -1 | cut -c1-200
