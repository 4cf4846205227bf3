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
    # general-geometry path (csrc/se_generic.cuh): power-of-two and Bluestein sizes, hops that do not divide n_fft,
    # rows shorter than n_fft, the DCCRN convention with its pinv parity correction
    for n, h, w, N in ((256, 64, 256, 1500), (512, 160, 400, 2400), (4096, 1024, 4096, 5000), (320, 160, 320, 3000),
                       (400, 100, 400, 2222), (64, 16, 64, 333), (512, 128, 512, 300)):
        c = types.SimpleNamespace(n_fft=n, hop_length=h, win_length=w, center=True)
        x = torch.randn(3, 1, N, device=dev, requires_grad=True)
        spec = se.stft_custom(x, c)
        m = torch.randn(*spec.shape, device=dev, requires_grad=True)
        se.istft_custom(se.apply_mask(spec, m, "C"), N, c).sum().backward()
        se.enhance(x.detach(), m, c, "E", True).square().mean().backward()
    for wl, inc, nf in ((320, 160, 512), (400, 100, 1024), (300, 75, 360)):
        gst, gist = se.ConvSTFT(wl, inc, nf, "hamming", "complex"), se.ConviSTFT(wl, inc, nf, None, "hamming", "complex")
        gs = gst(torch.randn(2, 1, 3000, device=dev)).requires_grad_(True)
        gist(gs).sum().backward()
        mre, mim = (torch.randn(2, nf // 2 + 1, gs.shape[-1], device=dev, requires_grad=True) for _ in range(2))
        gist.forward_masked(gs.detach(), mre, mim, "E").square().sum().backward()
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
    # DCCRN tail on its planar layout, and fused into ConviSTFT
    for mode in ("E", "C", "R"):
        mre, mim = (torch.randn(2, 257, s.shape[-1], device=dev, requires_grad=True) for _ in range(2))
        se.apply_mask_dccrn(s, mre, mim, mode).square().sum().backward()
        ist.forward_masked(s.detach(), mre, mim, mode).square().sum().backward()
    se.loss_phase_sensitive_spectral_approximation(spec.detach().clone().requires_grad_(True), spec.detach() * 0.5,
                                                   spec.detach() * 1.5).backward()
    # the peer-memory exchange kernel with a world of one (same code path: post, flag, poll, rank-ordered sum, value)
    import ctypes
    from speech_enhancement_pytorch_b200 import _native as nv
    L = nv.lib()
    local, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
    nv.check(L.se_p2p_create(ctypes.byref(local), handle))
    ptrs = (ctypes.c_void_p * 1)(local.value)
    sums = torch.rand(9, dtype=torch.float64, device=dev) + 1.0
    loss = torch.empty((), device=dev)
    for _ in range(3):
        nv.check(L.se_mrstft_exchange_value(sums.data_ptr(), ptrs, 1, 0, 3, 7000, loss.data_ptr(), nv.stream_ptr(torch.device(dev, 0))))
    torch.cuda.synchronize()
    nv.check(L.se_p2p_destroy(local))
    cfg = types.SimpleNamespace(dset=types.SimpleNamespace(norm="z-score", sample_rate=16000),
                                model=types.SimpleNamespace(name="dnn", segment=0.256, n_fft=512, hop_length=128,
                                                            win_length=512, center=True))
    se.evaluate(torch.randn(1, 2, 9000), None, dev, cfg)               # row_stats -> segment STFT (z-score fill) -> stitching iSTFT
    cfg.dset.norm = "none"
    se.evaluate(torch.randn(2, 1, 5000), None, dev, cfg)
    # round 2: 10-double exchange (row counts travel with the sums), DCCRN windows + polar ops, float64 inputs
    sums10 = torch.rand(10, dtype=torch.float64, device=dev) + 1.0
    nv.check(L.se_p2p_create(ctypes.byref(local), handle))
    ptrs = (ctypes.c_void_p * 1)(local.value)
    nv.check(L.se_mrstft_exchange_rows_value(sums10.data_ptr(), ptrs, 1, 0, 7000, loss.data_ptr(), nv.stream_ptr(torch.device(dev, 0))))
    nv.check(L.se_mrstft_loss_value_dev(sums10.data_ptr(), 7000, loss.data_ptr(), nv.stream_ptr(torch.device(dev, 0))))
    torch.cuda.synchronize()
    nv.check(L.se_p2p_destroy(local))
    st2, ist2 = se.ConvSTFT(400, 100, 512), se.ConviSTFT(400, 100, 512, 3000)          # hamming, feature_type='real'
    mags, phase = st2(torch.randn(2, 1, 3000, device=dev))
    mags.requires_grad_(True)
    phase.requires_grad_(True)
    ist2(mags, phase).sum().backward()
    c64 = types.SimpleNamespace(n_fft=512, hop_length=128, win_length=512, center=True)
    se.istft_custom(se.stft_custom(torch.randn(1, 1, 3000, device=dev, dtype=torch.float64), c64), 3000, c64)
    torch.cuda.synchronize()
    print("sanitizer smoke done")


if __name__ == "__main__":
    main()
