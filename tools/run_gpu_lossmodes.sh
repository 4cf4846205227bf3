#!/bin/bash
# A/B visit for the MR-STFT loss backward: saved-spectrum (default) vs recompute, launch order, n=2048 geometry
set -u
mkdir -p gpurun_out
echo "== pytest (loss)" ; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "mrstft or chain or cfg1 or graph or reentrant" > gpurun_out/pytest_loss.log 2>&1 ; echo "pytest rc=$?" ; tail -6 gpurun_out/pytest_loss.log
summ() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], "value", round(d['value']), "ms/step", round(d['ms_per_step'], 4), "loss", d['loss'])
for k in d['kernels']:
    if 'loss' in k['name'] or 'stft' in k['name']: print(f"  {k['name']:28s} {k['us']:8.1f} us  {k['gbs']:8.1f} GB/s  fp32 {k['tflops_fp32']}")
PY
}
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err || tail -3 gpurun_out/bench_$name.err
  summ gpurun_out/bench_$name.json
}
run default X=1
run recompute SE_MRSTFT_RECOMPUTE=1
run ascending SE_MRSTFT_BWD_ASCENDING=1
run fr8 SE_MRSTFT_BWD_FR8=1
run fr8_ascending SE_MRSTFT_BWD_FR8=1 SE_MRSTFT_BWD_ASCENDING=1
