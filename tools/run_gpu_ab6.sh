#!/bin/bash
# groups per CTA for the analysis-type kernels: default target (1184 CTAs -> 1 group per CTA at cfg2) vs fewer, longer CTAs
set -u
mkdir -p gpurun_out
for eng in 1 3; do for tgt in 0 512 296 256 148; do
  echo "== SE_ENGINE=$eng SE_TARGET_CTAS=$tgt"
  SE_ENGINE=$eng SE_TARGET_CTAS=$tgt timeout 600 python bench.py --steps 10 --warmup 3 --rounds 1 --no-cpu-baseline --no-e2e --no-configs --no-incumbent > gpurun_out/b.json 2> gpurun_out/b.err
  python - <<PY
import json
d=json.load(open('gpurun_out/b.json'))
k={x['name']:x['us'] for x in d['kernels']}
print("  stft", k['stft_fwd'], "istft_bwd", k['istft_bwd'], "loss_fwd", k['mrstft_loss_fwd(3 res)'], "enh_bwd", k['enhance_bwd'], "step", round(d['ms_per_step'],4))
PY
done; done
