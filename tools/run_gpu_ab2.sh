#!/bin/bash
# pair engine (default) vs scalar engine (SE_ENGINE=1): gpu tests on the default, bench on both
set -u
mkdir -p gpurun_out
echo "== pytest (pair engine)" ; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest.log 2>&1 ; echo "pytest rc=$?" ; tail -6 gpurun_out/pytest.log
for eng in 2 1; do
  echo "== bench SE_ENGINE=$eng"
  SE_ENGINE=$eng timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_eng$eng.json 2> gpurun_out/bench_eng$eng.err ; echo "bench rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_eng$eng.json'))
print("value", round(d['value']), "ms/step", round(d['ms_per_step'],4), "alts", [(a['composition'], round(a['ms_per_step'],4)) for a in d['alt_compositions']], "e2e", round(d['e2e']['value']))
for k in d['kernels']: print(f"  {k['name']:28s} {k['us']:8.1f} us  {k['gbs']:8.1f} GB/s  hbm {k.get('hbm_frac')}  fp32 {k['tflops_fp32']}")
PY
  tail -3 gpurun_out/bench_eng$eng.err
done
