"""Small invocation of every kernel family, meant to be run under compute-sanitizer on the GPU box:

    compute-sanitizer --tool memcheck  python tests/sanitizer_smoke.py
    compute-sanitizer --tool racecheck python tests/sanitizer_smoke.py
    compute-sanitizer --tool initcheck python tests/sanitizer_smoke.py
"""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_enhancement_pytorch_b200 as se  # noqa: E402


def main():
    torch.manual_seed(0)
    dev = "cuda"
    for n, h, N in ((512, 128, 5000), (1024, 256, 9000), (2048, 512, 20001), (1024, 512, 6000)):
        c = types.SimpleNamespace(n_fft=n, hop_length=h, win_length=n, center=True)
        x = torch.randn(2, 1, N, device=dev, requires_grad=True)
        spec = se.stft_custom(x, c)
        m = torch.randn(*spec.shape, device=dev, requires_grad=True)
        y = se.istft_custom(se.apply_mask(spec, m, "E", True), N, c)
        y.sum().backward()
        y2 = se.enhance(x.detach(), m, c, "C")
        y2.square().mean().backward()
        for mode in ("real", "E", "C", "R"):
            mm = torch.randn(*(spec.shape[:-1] if mode == "real" else spec.shape), device=dev, requires_grad=True)
            se.apply_mask_istft(spec.detach(), mm, N, c, mode, mode == "E").square().sum().backward()
        sp, ft = se.stft_custom_with_feature(x.detach(), c, "amplitude")
        tgt = torch.randn(2, 1, N, device=dev)
        e = spec.detach().clone().requires_grad_(True)
        se.loss_spectral(e, tgt, c, "mse").backward()
    est = torch.randn(3, 1, 7000, device=dev, requires_grad=True)
    ref = torch.randn(3, 1, 7000, device=dev)
    se.loss_mrstft(est, ref).backward()
    se.loss_sisdr(est, ref).backward()
    se.SI_SDR(ref, est.detach())
    frames = torch.randn(2, 3, 57, 40, device=dev, requires_grad=True)
    se.overlap_and_add(frames, 20).sum().backward()
    se.overlap_and_add(torch.randn(3, 9, 10, device=dev), 4)
    st, ist = se.ConvSTFT(400, 100, 512, "hann", "complex"), se.ConviSTFT(400, 100, 512, 3000, "hann", "complex")
    s = st(torch.randn(2, 1, 3000, device=dev)).requires_grad_(True)
    ist(s).sum().backward()
    cfg = types.SimpleNamespace(dset=types.SimpleNamespace(norm="z-score", sample_rate=16000),
                                model=types.SimpleNamespace(name="dnn", segment=0.256, n_fft=512, hop_length=128,
                                                            win_length=512, center=True))
    se.evaluate(torch.randn(1, 2, 9000), None, dev, cfg)
    torch.cuda.synchronize()
    print("sanitizer smoke done")


if __name__ == "__main__":
    main()
