#!/bin/bash
# A/B visit for the MR-STFT loss backward: recompute (default) vs the opt-in saved-spectrum mode.
# (The launch-order / cache-hint / geometry variants recorded in profiles/r01_notes.md (d) were measured with
# temporary switches that have since been folded into the saved-spectrum path or removed.)
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
run recompute X=1
run saved SE_MRSTFT_SAVE_SPECTRUM=1
run recompute2 X=1
run saved2 SE_MRSTFT_SAVE_SPECTRUM=1
