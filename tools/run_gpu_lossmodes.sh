#!/bin/bash
# A/B visit for the MR-STFT loss backward: saved-spectrum (default) vs recompute, launch order, cache hints
set -u
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
ks = {k['name']: k['us'] for k in d['kernels']}
print(f"{sys.argv[1]:44s} step {d['ms_per_step']*1e3:7.1f} us  loss_fwd {ks.get('mrstft_loss_fwd(3 res)')}  loss_bwd {ks.get('mrstft_loss_bwd(3 res)')}  tail_bwd {ks.get('mask_istft_bwd')}")
PY
}
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err || tail -3 gpurun_out/bench_$name.err
  summ gpurun_out/bench_$name.json
}
run recompute SE_MRSTFT_RECOMPUTE=1
run saved X=1
run saved_asc SE_MRSTFT_BWD_ASCENDING=1
run saved_asc_h1 SE_MRSTFT_BWD_ASCENDING=1 SE_MRSTFT_HINTS=1
run saved_asc_h2 SE_MRSTFT_BWD_ASCENDING=1 SE_MRSTFT_HINTS=2
run saved_asc_h3 SE_MRSTFT_BWD_ASCENDING=1 SE_MRSTFT_HINTS=3
run saved_asc_h4 SE_MRSTFT_BWD_ASCENDING=1 SE_MRSTFT_HINTS=4
run saved_asc_h7 SE_MRSTFT_BWD_ASCENDING=1 SE_MRSTFT_HINTS=7
run saved_h7 SE_MRSTFT_HINTS=7
run recompute2 SE_MRSTFT_RECOMPUTE=1
