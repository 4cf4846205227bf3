#!/bin/bash
# quick GPU visit: gpu tests + bench (+ optional launch list)
set -u
mkdir -p gpurun_out
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest.log 2>&1 ; echo "pytest rc=$?" ; tail -12 gpurun_out/pytest.log
echo "== bench" ; timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "bench rc=$?" ; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print("value", round(d['value']), "ms/step", round(d['ms_per_step'],4), "comp", d['composition'], "alts", [(a['composition'], round(a['ms_per_step'],4)) for a in d['alt_compositions']], "e2e", round(d['e2e']['value']), d['e2e']['ms_per_step'])
for k in d['kernels']: print(f"  {k['name']:28s} {k['us']:8.1f} us  {k['gbs']:8.1f} GB/s  hbm {k.get('hbm_frac')}  fp32 {k['tflops_fp32']}")
print(d['roofline'])
PY
tail -3 gpurun_out/bench.err
