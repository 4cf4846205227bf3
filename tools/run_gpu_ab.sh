#!/bin/bash
# A/B of compile-time kernel variants in ONE GPU visit: every speech_enhancement_pytorch_b200/libse_variant_*.so
# (complete libraries built with an extra -D flag) is swapped in for libse_b200.so and benched, interleaved with the
# shipped build (run-to-run spread of the step is ~0.3 us on one box, box-to-box a few us).
summ() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
ks = {k['name']: k['us'] for k in d['kernels']}
print(f"{sys.argv[1]:48s} step {d['ms_per_step']*1e3:7.1f} us  loss {d['loss']:.7f} stft {ks.get('stft_fwd')} tail {ks.get('mask_istft_fwd')}/{ks.get('mask_istft_bwd')} loss_fwd {ks.get('mrstft_loss_fwd(3 res)')}  loss_bwd {ks.get('mrstft_loss_bwd(3 res)')}")
PY
}
mkdir -p gpurun_out
L=speech_enhancement_pytorch_b200
cp $L/libse_b200.so /tmp/base.so
for rep in 1 2; do
  for lib in /tmp/base.so $L/libse_variant_*.so; do
    tag=$(basename $lib .so)
    cp $lib $L/libse_b200.so
    timeout 300 python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_ab_$tag.json 2>/dev/null; summ gpurun_out/bench_ab_$tag.json
  done
done
cp /tmp/base.so $L/libse_b200.so
