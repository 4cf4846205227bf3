#!/usr/bin/env python
"""Kernel-level throughput of every BASELINE.json config shape on one GPU (device-resident inputs, CUDA events,
inputs rotated over 4 buffer sets).  Prints one JSON object; `bench.py` remains the contract benchmark (cfg 2)."""
import json
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_enhancement_pytorch_b200 as se  # noqa: E402

HBM = 6534.8


def timed(fn, sets, reps=30, warm=5):
    for i in range(warm):
        fn(sets[i % len(sets)])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(sets[i % len(sets)])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def cfg(n, h):
    return types.SimpleNamespace(n_fft=n, hop_length=h, win_length=n, center=True)


def main():
    out = {}
    dev = "cuda"
    # cfg 1: 16 x 4 s, 512/128, real mask (Unet), inference
    c = cfg(512, 128)
    sets = [(torch.randn(16, 1, 64000, device=dev), torch.rand(16, 1, 257, 501, device=dev)) for _ in range(4)]
    with torch.no_grad():
        us = timed(lambda s: se.istft_custom(se.apply_mask(se.stft_custom(s[0], c), s[1], "real"), 64000, c), sets)
        uf = timed(lambda s: se.enhance(s[0], s[1], c, "real"), sets)
    S, P, M = 4 * 64000, 8 * 257 * 501, 4 * 257 * 501
    out["cfg1_stft_mask_istft"] = {"us": round(us, 1), "audio_s_per_s": round(64 / us * 1e6), "alg_GBs": round(16 * (2 * S + 4 * P + M) / us * 1e-3)}
    out["cfg1_fused_enhance"] = {"us": round(uf, 1), "audio_s_per_s": round(64 / uf * 1e6), "alg_GBs": round(16 * (2 * S + M) / uf * 1e-3),
                                 "hbm_frac": round(16 * (2 * S + M) / uf * 1e-3 / HBM, 3)}
    # cfg 3: MR-STFT loss fwd+bwd, 128 x 4 s
    sets = [(torch.randn(128, 1, 64000, device=dev, requires_grad=True), torch.randn(128, 1, 64000, device=dev)) for _ in range(4)]

    def loss_step(s):
        s[0].grad = None
        se.loss_mrstft(s[0], s[1]).backward()
    us = timed(loss_step, sets)
    out["cfg3_mrstft_fwd_bwd_128x4s"] = {"us": round(us, 1), "audio_s_per_s": round(512 / us * 1e6),
                                         "tflops_fp32": round(128 * 5 * sum(2.5 * n * (n.bit_length() - 1) * (1 + 64000 // (n // 4)) for n in (512, 1024, 2048)) / us * 1e-6, 2)}
    # cfg 4: DCCRN transforms 16 x 4 s, wav -> spec -> wav with a polar mask in between
    st, ist = se.ConvSTFT(400, 100, 512, "hann", "complex"), se.ConviSTFT(400, 100, 512, 64000, "hann", "complex")
    sets = [(torch.randn(16, 1, 64000, device=dev), torch.randn(16, 257, 643, device=dev), torch.randn(16, 257, 643, device=dev)) for _ in range(4)]
    with torch.no_grad():
        us = timed(lambda s: ist(se.apply_mask_dccrn(st(s[0]), s[1], s[2], "E")), sets)
    out["cfg4_dccrn_transforms_16x4s"] = {"us": round(us, 1), "audio_s_per_s": round(64 / us * 1e6)}
    with torch.no_grad():
        us = timed(lambda s: ist.forward_masked(st(s[0]), s[1], s[2], "E"), sets)
    out["cfg4_dccrn_fused_tail_16x4s"] = {"us": round(us, 1), "audio_s_per_s": round(64 / us * 1e6)}
    # cfg 5: 44.1 kHz stereo 30 s clips, 8 clips, n_fft 2048 and 1024, complex mask, fused and unfused
    for n in (2048, 1024):
        c = cfg(n, n // 4)
        F, T = n // 2 + 1, 1 + 1323000 // (n // 4)
        sets = [(torch.randn(8, 2, 1323000, device=dev), torch.rand(8, 2, F, T, 2, device=dev) * 2 - 1) for _ in range(2)]
        with torch.no_grad():
            uf = timed(lambda s: se.enhance(s[0], s[1], c, "C"), sets, reps=10, warm=3)
            uu = timed(lambda s: se.istft_custom(se.apply_mask(se.stft_custom(s[0], c), s[1], "C"), 1323000, c), sets, reps=10, warm=3)
        S, Mc = 4 * 1323000, 8 * F * T
        out[f"cfg5_fused_enhance_n{n}"] = {"us": round(uf, 1), "audio_s_per_s": round(240 / uf * 1e6), "channel_s_per_s": round(480 / uf * 1e6),
                                           "alg_GBs": round(16 * (2 * S + Mc) / uf * 1e-3), "hbm_frac": round(16 * (2 * S + Mc) / uf * 1e-3 / HBM, 3)}
        out[f"cfg5_unfused_n{n}"] = {"us": round(uu, 1), "audio_s_per_s": round(240 / uu * 1e6)}
        del sets
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
