"""evaluate() on a 30 s 16 kHz stereo clip (the reference-faithful cfg5 variant, SURVEY 8d: 814 segments x 4 s, stride 512):
round-2 flow (row_stats -> segment_stft with folded z-score -> istft_stitch) against the round-1 flow rebuilt from the
same public functions (torch z-score, segment_stft, istft_custom of every sample of every segment, torch stitch)."""
import json
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_enhancement_pytorch_b200 as se  # noqa: E402
ev = sys.modules["speech_enhancement_pytorch_b200.evaluate"]      # the package re-exports the function under the same name


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def r1_flow(x, conf):
    with torch.no_grad():
        mean = torch.mean(x, dim=-1, keepdim=True)
        std = torch.std(x, dim=-1, keepdim=True)
        xn = (x - mean) / (std + 1e-9)
        stride, nfeat = conf.model.win_length, int(conf.dset.sample_rate * conf.model.segment)
        batch, nseg = ev.segment_stft(xn, nfeat, stride, conf.model)
        out = se.istft_custom(torch.unsqueeze(batch, 1), nfeat, conf.model)
        out = out.reshape(nseg, x.shape[0], x.shape[1], nfeat)
        return ev.stitch_segments(out, nfeat, stride, x.shape[-1]) * (std + 1e-9) + mean


def main():
    conf = types.SimpleNamespace(dset=types.SimpleNamespace(norm="z-score", sample_rate=16000),
                                 model=types.SimpleNamespace(name="unet", segment=4.0, n_fft=512, hop_length=128, win_length=512,
                                                             center=True, sources=["clean"]))
    x = (0.3 * torch.randn(1, 2, 480000) + 0.01).cuda()
    a = se.evaluate(x, None, "cuda", conf)
    b = r1_flow(x, conf)
    out = {"clip": "30 s, 16 kHz, stereo; 4 s segments, stride 512 -> 814 segments x 2 channels",
           "max_abs_diff_r2_vs_r1_flow": float((a - b).abs().max()), "identity_err": float((a - x).abs().max()),
           "r2_ms": timed(lambda: se.evaluate(x, None, "cuda", conf)), "r1_flow_ms": timed(lambda: r1_flow(x, conf))}
    out["speedup"] = out["r1_flow_ms"] / out["r2_ms"]
    out["audio_s_per_s_r2"] = 30.0 / (out["r2_ms"] * 1e-3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
