#!/bin/bash
# N-GPU visit: 2-rank exchange tests (peer-memory kernel and NCCL), then the bench with each exchange flavour
set -u
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
if [ "$N" = "2" ]; then echo "== pytest distributed" ; timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_dist.log 2>&1 ; echo "rc=$?" ; tail -5 gpurun_out/pytest_dist.log; fi
run() { # tag, env
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --no-breakdown > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err ; echo "bench N=$N $tag rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_n${N}_$tag.json'))
print("N", d['n_gpus'], "$tag", "value", round(d['value']), "ms/step", round(d['ms_per_step'],4), "e2e", round(d['e2e']['value']), "loss", d['loss'], "|", d['config'].get('exchange'))
PY
  tail -3 gpurun_out/bench_n${N}_$tag.err
}
run p2p SE_P2P_EXCHANGE=1
run nccl SE_P2P_EXCHANGE=0

