"""Times the GPU incumbent: the reference's own torch calls (oracle restatement) on CUDA tensors -- cuFFT /
ATen elementwise / autograd -- against this repo's kernels, same cfg2 step as bench.py.  Measurement tool,
not a test.  Run on the GPU box:  python tests/time_gpu_incumbent.py"""
import json
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_enhancement_pytorch_b200 as se  # noqa: E402
from oracle import spectral_oracle as oref  # noqa: E402


def timed(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def main():
    rows, N = 64, 64000
    c = types.SimpleNamespace(n_fft=1024, hop_length=256, win_length=1024, center=True)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(rows, 1, N, generator=g).cuda()
    clean = (x.cpu() + 0.3 * torch.randn(rows, 1, N, generator=g)).cuda()
    raw = torch.randn(rows, 1, 513, 251, 2, generator=g).cuda().requires_grad_(True)

    def step(stft, mask, istft, loss):
        raw.grad = None
        y = istft(mask(stft(x, c), raw), N, c)
        loss(y, clean).backward()

    out = {}
    out["step_ours_autograd_us"] = timed(lambda: step(se.stft_custom, lambda s, m: se.apply_mask(s, m, "E", True),
                                                      se.istft_custom, se.loss_mrstft))
    out["step_incumbent_us"] = timed(lambda: step(oref.stft_custom_ref, lambda s, m: oref.mask_apply_ref(s, m, "E", True),
                                                  oref.istft_custom_ref, oref.mrstft_loss_ref), reps=10, warm=3)
    with torch.no_grad():
        spec = se.stft_custom(x, c)
        out["stft_ours_us"] = timed(lambda: se.stft_custom(x, c))
        out["stft_incumbent_us"] = timed(lambda: oref.stft_custom_ref(x, c))
        out["istft_ours_us"] = timed(lambda: se.istft_custom(spec, N, c))
        out["istft_incumbent_us"] = timed(lambda: oref.istft_custom_ref(spec, N, c))
        out["mask_ours_us"] = timed(lambda: se.apply_mask(spec, raw.detach(), "E", True))
        out["mask_incumbent_us"] = timed(lambda: oref.mask_apply_ref(spec, raw.detach(), "E", True))
        out["mrstft_fwd_ours_us"] = timed(lambda: se.loss_mrstft(x, clean))
        out["mrstft_fwd_incumbent_us"] = timed(lambda: oref.mrstft_loss_ref(x, clean), reps=10, warm=3)
    print(json.dumps({k: round(v, 1) for k, v in out.items()}))


if __name__ == "__main__":
    main()
