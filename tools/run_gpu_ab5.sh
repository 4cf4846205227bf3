#!/bin/bash
# three-pass scalar engine (default) vs two-pass engine (SE_ENGINE=3)
set -u
mkdir -p gpurun_out
for cfg in "SE_ENGINE=1" "SE_ENGINE=3"; do
  echo "== bench $cfg"
  env $cfg timeout 600 python bench.py --steps 30 --warmup 5 --rounds 3 --no-cpu-baseline --no-e2e --no-configs --no-incumbent > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err ; echo "bench rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$cfg.json'))
print("value", round(d['value']), "ms/step", round(d['ms_per_step'],4), "alts", [(a['composition'], round(a['ms_per_step'],4)) for a in d['alt_compositions']])
for k in d['kernels']: print(f"  {k['name']:28s} {k['us']:8.1f} us  hbm {k.get('hbm_frac')}  fp32 {k['tflops_fp32']}")
PY
  tail -2 gpurun_out/bench_$cfg.err
done
