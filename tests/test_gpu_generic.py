"""General-geometry transforms (any power-of-two n_fft, any hop / win_length; any DCCRN win_len / win_inc / fft_len) on the
GPU, through the public API, against goldens produced by running the real reference (tests/golden/general_geometry.npz,
make_golden.py::gen_general_geometry) and against the torch oracle on seeded inputs.  Tolerance 1e-4 (north_star)."""
import types

import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def se():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import speech_enhancement_pytorch_b200 as m
    m._native.lib()
    return m


@pytest.fixture(scope="module")
def oref():
    from oracle import spectral_oracle
    return spectral_oracle


def cfg(n, h, w):
    return types.SimpleNamespace(n_fft=n, hop_length=h, win_length=w, center=True)


def rel(a, b):
    a = a.detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_reference_run_goldens_torch_convention(se):
    g = golden("general_geometry")
    for i, (N, n, h, w) in enumerate(g["t_meta"]):
        c = cfg(int(n), int(h), int(w))
        k = lambda name: torch.from_numpy(g[f"t{i}_{name}"]).cuda()
        x = k("x").requires_grad_(True)
        spec = se.stft_custom(x, c)
        assert rel(spec, k("spec")) < TOL, (i, "stft")
        (gx,) = torch.autograd.grad(spec, x, k("gspec"))
        assert rel(gx, k("gx")) < TOL, (i, "stft adjoint")
        s = k("s").requires_grad_(True)
        y = se.istft_custom(s, int(N), c)
        assert rel(y, k("y")) < TOL, (i, "istft")
        (gs,) = torch.autograd.grad(y, s, k("gy"))
        assert rel(gs, k("gs")) < TOL, (i, "istft adjoint")


def test_reference_run_goldens_without_centre_padding(se):
    g = golden("general_geometry")
    for i, (N, n, h, w) in enumerate(g["n_meta"]):
        c = types.SimpleNamespace(n_fft=int(n), hop_length=int(h), win_length=int(w), center=False)
        k = lambda name: torch.from_numpy(g[f"n{i}_{name}"]).cuda()
        x = k("x").requires_grad_(True)
        spec = se.stft_custom(x, c)
        assert spec.shape == k("spec").shape
        assert rel(spec, k("spec")) < TOL
        (gx,) = torch.autograd.grad(spec, x, k("gspec"))
        assert rel(gx, k("gx")) < TOL


def test_reference_run_goldens_dccrn_convention(se):
    g = golden("general_geometry")
    for i, (N, wl, inc, nfft, wt) in enumerate(g["c_meta"]):
        wtype = "hamming" if wt else "hann"
        st = se.ConvSTFT(int(wl), int(inc), int(nfft), wtype, "complex").cuda()
        ist = se.ConviSTFT(int(wl), int(inc), int(nfft), None, wtype, "complex").cuda()
        k = lambda name: torch.from_numpy(g[f"c{i}_{name}"]).cuda()
        assert rel(st(k("x")), k("spec")) < TOL, (i, "ConvSTFT")
        s = k("s").requires_grad_(True)
        y = ist(s)
        assert y.shape == k("y").shape
        assert rel(y, k("y")) < TOL, (i, "ConviSTFT")
        (gs,) = torch.autograd.grad(y, s, k("gy"))
        assert rel(gs, k("gs")) < TOL, (i, "ConviSTFT adjoint")


@pytest.mark.parametrize("n,hop,win,N,rows", [(256, 64, 256, 16000, 8), (512, 160, 400, 16000, 8), (4096, 1024, 4096, 40000, 3),
                                              (8192, 2048, 8192, 50000, 2), (1024, 100, 1024, 9000, 4), (128, 32, 128, 5000, 5),
                                              (512, 128, 512, 300, 3), (320, 160, 320, 16000, 8), (400, 100, 400, 16000, 8), (1000, 250, 800, 9000, 3)])
def test_seeded_vs_oracle(se, oref, n, hop, win, N, rows):
    c = cfg(n, hop, win)
    x = torch.randn(rows, 1, N, generator=torch.Generator().manual_seed(n + hop))
    want = oref.stft_custom_ref(x, c)
    xg = x.cuda().requires_grad_(True)
    spec = se.stft_custom(xg, c)
    assert rel(spec, want) < TOL
    y = se.istft_custom(spec, N, c)
    assert rel(y, oref.istft_custom_ref(want, N, c)) < TOL
    assert rel(y, x) < TOL                                    # round trip
    # gradient through both transforms vs float64 autograd of the oracle
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(5))
    (gx,) = torch.autograd.grad(y, xg, gy.cuda())
    x64 = x.double().requires_grad_(True)
    y64 = oref.istft_custom_ref(oref.stft_custom_ref(x64, c), N, c)
    (g64,) = torch.autograd.grad(y64, x64, gy.double())
    assert rel(gx, g64) < 1e-3


def test_enhance_composes_for_general_geometry(se, oref):
    """se.enhance / se.apply_mask_istft have fused kernels for the tuned geometries only; elsewhere they run the three
    (two) stages -- same result, same gradient."""
    c = cfg(256, 64, 256)
    N = 4000
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(3, 1, N, generator=gen)
    mask = torch.randn(3, 1, 129, 1 + N // 64, 2, generator=gen)
    m64 = mask.double().requires_grad_(True)
    spec64 = oref.stft_custom_ref(x.double(), c)
    y64 = oref.istft_custom_ref(oref.mask_apply_ref(spec64, m64, "E", pre_tanh=True), N, c)
    gy = torch.randn(y64.shape, generator=gen)
    (g64,) = torch.autograd.grad(y64, m64, gy.double())
    mg = mask.cuda().requires_grad_(True)
    y = se.enhance(x.cuda(), mg, c, "E", pre_tanh=True)
    assert rel(y, y64) < TOL
    (gm,) = torch.autograd.grad(y, mg, gy.cuda())
    assert rel(gm, g64) < 1e-3
    y2 = se.apply_mask_istft(se.stft_custom(x.cuda(), c), mask.cuda(), N, c, "E", pre_tanh=True)
    assert rel(y2, y64) < TOL


def test_dccrn_general_geometry_tail_and_errors(se, oref):
    st, ist = se.ConvSTFT(320, 160, 512, "hann", "complex").cuda(), se.ConviSTFT(320, 160, 512, None, "hann", "complex").cuda()
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(2, 1, 3200, generator=gen).cuda()
    spec = st(x)
    mre = torch.randn(2, 257, spec.shape[-1], generator=gen).cuda().requires_grad_(True)
    mim = torch.randn(2, 257, spec.shape[-1], generator=gen).cuda().requires_grad_(True)
    y = ist.forward_masked(spec, mre, mim, "E")
    sp = spec.cpu().double()
    masked = oref.mask_apply_ref(torch.stack([sp[:, :257], sp[:, 257:]], -1),
                                 torch.stack([mre.detach().cpu().double(), mim.detach().cpu().double()], -1), "E")
    want = oref.conv_istft_ref(torch.cat([masked[..., 0], masked[..., 1]], 1), 320, 160, 512, "hann", None)
    assert rel(y, want) < TOL
    y.sum().backward()
    assert torch.isfinite(mre.grad).all() and torch.isfinite(mim.grad).all()
    with pytest.raises((NotImplementedError, RuntimeError)):
        se.stft_custom(x, cfg(401, 100, 401))                # odd n_fft
    with pytest.raises(RuntimeError, match="overlap add"):
        se.istft_custom(torch.zeros(1, 1, 129, 5, 2).cuda(), 900, cfg(256, 256, 256))


def test_general_geometry_large_batch_timing_sanity(se):
    """64 x 4 s at 256/64: the general path stays within an order of magnitude of the tuned engine's rate (it is the
    compatibility path, not the product's fast path) and handles a full-size batch."""
    c = cfg(256, 64, 256)
    x = torch.randn(64, 1, 64000).cuda()
    spec = se.stft_custom(x, c)
    y = se.istft_custom(spec, 64000, c)
    assert rel(y, x) < TOL


@pytest.mark.parametrize("model", [None, "mask"])
def test_evaluate_at_a_general_geometry(se, oref, model):
    """evaluate() (src/evaluate.py:10-98) with the commented CRN setting n_fft 320 / hop 160 (config.yaml:78-80): no fused
    segment / stitch kernels there, the plain general-geometry transforms run around the model."""
    g = torch.Generator().manual_seed(21)
    mix = torch.randn(1, 2, 30000, generator=g) * 0.3 + 0.05
    config = types.SimpleNamespace(
        dset=types.SimpleNamespace(norm="z-score", sample_rate=16000),
        model=types.SimpleNamespace(name="crn", segment=0.5, n_fft=320, hop_length=160, win_length=320, center=True))
    fn = None
    if model:
        class Toy(torch.nn.Module):
            def forward(self, spec):
                f = torch.linspace(0.2, 1.0, spec.shape[-3], device=spec.device)[:, None, None]
                return spec * f
        fn = Toy()
    want = oref.evaluate_ref(mix, fn, config)
    got = se.evaluate(mix, fn.cuda() if fn else None, "cuda", config)
    assert got.shape == want.shape
    assert rel(got, want) < TOL


@pytest.mark.parametrize("n,hop,win", [(512, 128, 512), (1024, 256, 1024), (320, 160, 320)])
def test_stft_custom_without_centre_padding(se, oref, n, hop, win):
    """config.center = False: the reference's stft_custom passes it to torch.stft (src/evaluate.py:116)."""
    c = types.SimpleNamespace(n_fft=n, hop_length=hop, win_length=win, center=False)
    x = torch.randn(3, 1, 9000, generator=torch.Generator().manual_seed(n))
    x64 = x.double().requires_grad_(True)
    want = oref.stft_custom_ref(x64, c)
    xg = x.cuda().requires_grad_(True)
    got = se.stft_custom(xg, c)
    assert got.shape == want.shape
    assert rel(got, want) < TOL
    g = torch.randn(want.shape, generator=torch.Generator().manual_seed(1))
    (g64,) = torch.autograd.grad(want, x64, g.double())
    (gx,) = torch.autograd.grad(got, xg, g.cuda())
    assert rel(gx, g64) < TOL
    with pytest.raises(RuntimeError, match="overlap add"):
        se.istft_custom(got, 9000, c)


def test_feature_epilogue_and_tuned_only_ops_at_a_general_geometry(se, oref):
    c = cfg(256, 64, 256)
    x = torch.randn(2, 1, 3000, generator=torch.Generator().manual_seed(4))
    spec, feat = se.stft_custom_with_feature(x.cuda(), c, "magnitude")
    want = oref.stft_custom_ref(x, c)
    assert rel(spec, want) < TOL
    assert rel(feat, oref.magnitude_feature_ref(want, "magnitude")) < TOL
    # fused kernels that only exist for the tuned geometries say so instead of running something else
    with pytest.raises(NotImplementedError, match="512/1024/2048"):
        se.loss_spectral(spec, x.cuda(), c, "mse")
