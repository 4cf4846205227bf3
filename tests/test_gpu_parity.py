"""Parity of the CUDA path (through the C-ABI) against the oracle and the reference's golden
vectors.  Tolerances are BASELINE.json's: spectra / waveforms 1e-4 relative, loss / gradients
1e-3 relative (relative = max |err| / max |ref|)."""
import math
import types

import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu

TOL_SPEC = 1e-4
TOL_GRAD = 1e-3


@pytest.fixture(scope="module")
def se():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import speech_enhancement_pytorch_b200 as m
    m._native.lib()          # fail loudly if the extension is missing
    return m


@pytest.fixture(scope="module")
def oref():
    from oracle import spectral_oracle
    return spectral_oracle


def cfg(n, h, w):
    return types.SimpleNamespace(n_fft=n, hop_length=h, win_length=w, center=True)


def rel(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def directional_check(grad, g64, g_ref32):
    """ours vs float64 along the gradient and along random directions; the bar is 1e-3 or twice the
    deviation of the reference's own fp32 path, whichever is larger (the loss is ill-conditioned)."""
    g = grad.detach().cpu().double()
    nrm = float(g64.norm())
    along = float((g * g64).sum()) / nrm
    along_ref = float((g_ref32.detach().double() * g64).sum()) / nrm
    assert abs(along - nrm) < max(TOL_GRAD * nrm, 2.0 * abs(along_ref - nrm)), (along, along_ref, nrm)
    gen = torch.Generator().manual_seed(77)
    for _ in range(4):
        v = torch.randn(g64.shape, generator=gen).double()
        bound = 6.0 * (2 * TOL_GRAD * nrm) * float(v.norm()) / g64.numel() ** 0.5
        assert abs(float(((g - g64) * v).sum())) < bound


GOLD = ["stft_n512", "stft_n1024", "stft_n2048", "stft_n512_hop256", "stft_n512_win400", "stft_n512_4d",
        "stft_n512_short_len", "stft_n512_long_len", "stft_n1024_multiple", "stft_structured"]


@pytest.mark.parametrize("name", GOLD)
def test_golden_stft_istft(se, name):
    g = golden(name)
    n, h, w, length = (int(v) for v in g["meta"])
    c = cfg(n, h, w)
    x = torch.from_numpy(g["x"]).cuda()
    spec = se.stft_custom(x, c)
    assert spec.shape == g["spec"].shape and spec.dtype == torch.float32 and spec.is_contiguous()
    assert rel(spec, torch.from_numpy(g["spec"])) < TOL_SPEC
    y = se.istft_custom(torch.from_numpy(g["spec"]).cuda(), length, c)
    assert y.shape == g["y"].shape
    assert rel(y, torch.from_numpy(g["y"])) < TOL_SPEC
    if "spec2" in g:
        y2 = se.istft_custom(torch.from_numpy(g["spec2"]).cuda(), length, c)
        assert rel(y2, torch.from_numpy(g["y2"])) < TOL_SPEC


@pytest.mark.parametrize("name", ["grad_n512", "grad_n1024", "grad_n512_win400"])
def test_golden_autograd(se, name):
    g = golden(name)
    n, h, w, N = (int(v) for v in g["meta"])
    c = cfg(n, h, w)
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    spec = se.stft_custom(x, c)
    (gx,) = torch.autograd.grad(spec, x, torch.from_numpy(g["gspec"]).cuda())
    assert rel(gx, torch.from_numpy(g["gx"])) < TOL_SPEC
    s = torch.from_numpy(g["s"]).cuda().requires_grad_(True)
    y = se.istft_custom(s, N, c)
    (gs,) = torch.autograd.grad(y, s, torch.from_numpy(g["gy"]).cuda())
    assert rel(gs, torch.from_numpy(g["gs"])) < TOL_SPEC


@pytest.mark.parametrize("n,h,w,shape", [(512, 128, 512, (3, 2, 16000)), (1024, 256, 1024, (4, 1, 16384)),
                                         (2048, 512, 2048, (2, 1, 44100)), (1024, 512, 1024, (2, 2, 1, 9999)),
                                         (512, 128, 400, (2, 1, 16001))])
def test_seeded_vs_oracle(se, oref, n, h, w, shape):
    g = torch.Generator().manual_seed(n + shape[-1])
    x = torch.randn(*shape, generator=g)
    c = cfg(n, h, w)
    ref_spec = oref.stft_custom_ref(x, c)
    spec = se.stft_custom(x.cuda(), c)
    assert rel(spec, ref_spec) < TOL_SPEC
    pert = ref_spec + 0.01 * torch.randn(ref_spec.shape, generator=g)
    for length in (shape[-1], shape[-1] - 100):
        assert rel(se.istft_custom(pert.cuda(), length, c), oref.istft_custom_ref(pert, length, c)) < TOL_SPEC


def test_round_trip_property_full_size(se):
    """The reference's own pin: istft(stft(x)) == x to 1e-5 (test/test_train.py:97-100), at
    BASELINE config sizes (16x4 s / 64x4 s / 128x4 s)."""
    for n, h, rows in ((512, 128, 16), (1024, 256, 64), (2048, 512, 128)):
        x = torch.randn(rows, 1, 64000, device="cuda")
        c = cfg(n, h, n)
        y = se.istft_custom(se.stft_custom(x, c), 64000, c)
        assert float((y - x).abs().max()) < 1e-5


def test_linearity_and_adjoint_property_full_size(se):
    """<stft(x), G> == <x, stft^T(G)> and linearity, at cfg2 size."""
    c = cfg(1024, 256, 1024)
    x = torch.randn(64, 1, 64000, device="cuda", requires_grad=True)
    spec = se.stft_custom(x, c)
    G = torch.randn_like(spec)
    (gx,) = torch.autograd.grad(spec, x, G)
    lhs = float((spec.double() * G.double()).sum())
    rhs = float((x.double() * gx.double()).sum())
    assert abs(lhs - rhs) / abs(lhs) < 1e-5
    x2 = torch.randn_like(x)
    lin = se.stft_custom(x.detach() + 2 * x2, c) - spec.detach() - 2 * se.stft_custom(x2, c)
    assert float(lin.abs().max()) < 1e-5


@pytest.mark.parametrize("mode", ["real", "E", "C", "R"])
@pytest.mark.parametrize("pre_tanh", [False, True])
def test_masks_vs_oracle(se, oref, mode, pre_tanh):
    g = torch.Generator().manual_seed(5)
    spec = torch.randn(3, 1, 257, 41, 2, generator=g)
    spec[0, 0, 0, 0] = 0.0
    mask = torch.randn(*(spec.shape[:-1] if mode == "real" else spec.shape), generator=g)
    sc = spec.cuda().requires_grad_(True)
    mc = mask.cuda().requires_grad_(True)
    out = se.apply_mask(sc, mc, mode, pre_tanh)
    sr = spec.double().requires_grad_(True)
    mr = mask.double().requires_grad_(True)
    ref = oref.mask_apply_ref(sr, mr, mode, pre_tanh)
    assert rel(out, ref) < TOL_SPEC
    go = torch.randn(ref.shape, generator=g)
    go[0, 0, 0, 0] = 0.0
    gs, gm = torch.autograd.grad(out, (sc, mc), go.cuda())
    gsr, gmr = torch.autograd.grad(ref, (sr, mr), go.double())
    assert rel(gm, gmr) < TOL_GRAD
    assert rel(gs, torch.nan_to_num(gsr)) < TOL_GRAD


def test_mask_golden_from_reference_models(se):
    g = golden("dcunet_mask_E")
    out = se.apply_mask(torch.from_numpy(g["spec"]).cuda(), torch.from_numpy(g["raw_mask"]).cuda(), "E", True)
    assert rel(out, torch.from_numpy(g["out"])) < TOL_SPEC
    for mode in ("E", "C", "R"):
        g = golden(f"dccrn_mask_{mode}")
        out = se.apply_mask_dccrn(torch.from_numpy(g["specs"]).cuda(), torch.from_numpy(g["mask_re"]).cuda(),
                                  torch.from_numpy(g["mask_im"]).cuda(), mode)
        assert rel(out, torch.from_numpy(g["out_spec"])) < TOL_SPEC


@pytest.mark.parametrize("mode", ["E", "C", "R"])
def test_dccrn_planar_mask_matches_interleaved_and_autograd(se, oref, mode):
    """apply_mask_dccrn (one planar kernel each way) against the interleaved kernels composed with stack / cat, and
    against the oracle's autograd, at DCCRN's real shape [B,514,643]."""
    from speech_enhancement_pytorch_b200 import ops
    g = torch.Generator().manual_seed(31)
    specs = torch.randn(3, 514, 643, generator=g)
    mre, mim = torch.randn(3, 257, 643, generator=g), torch.randn(3, 257, 643, generator=g)
    go = torch.randn(3, 514, 643, generator=g)
    a = [t.cuda().requires_grad_(True) for t in (specs, mre, mim)]
    out = se.apply_mask_dccrn(a[0], a[1], a[2], mode)
    ga = torch.autograd.grad(out, a, go.cuda())
    b = [t.cuda().requires_grad_(True) for t in (specs, mre, mim)]
    spec_i = torch.stack([b[0][:, :257], b[0][:, 257:]], -1)
    o2 = ops.mask_apply(spec_i, torch.stack([b[1], b[2]], -1), mode, False)
    out2 = torch.cat([o2[..., 0], o2[..., 1]], 1)
    gb = torch.autograd.grad(out2, b, go.cuda())
    assert rel(out, out2) < 1e-6              # same per-bin math; FMA contraction may differ between the two kernels
    for x, y in zip(ga, gb):
        assert rel(x, y) < 1e-6
    c = [t.double().requires_grad_(True) for t in (specs, mre, mim)]
    o3 = oref.mask_apply_ref(torch.stack([c[0][:, :257], c[0][:, 257:]], -1), torch.stack([c[1], c[2]], -1), mode, False)
    out3 = torch.cat([o3[..., 0], o3[..., 1]], 1)
    gc = torch.autograd.grad(out3, c, go.double())
    assert rel(out, out3) < TOL_SPEC
    for x, y in zip(ga, gc):
        assert rel(x, y) < TOL_GRAD


@pytest.mark.parametrize("mode", ["E", "C", "R"])
def test_dccrn_fused_tail_matches_two_stage_and_oracle(se, oref, mode):
    """ConviSTFT.forward_masked (mask tail + ConviSTFT in one launch each way) against istft(apply_mask_dccrn(...)) and,
    through autograd, against the oracle's float64 composition; DCCRN's real shape."""
    g = torch.Generator().manual_seed(32)
    x = torch.randn(3, 1, 64000, generator=g)
    mre, mim = torch.randn(3, 257, 643, generator=g), torch.randn(3, 257, 643, generator=g)
    mre[:, 0] = 0.0
    mim[:, 0] = 0.0
    gy = torch.randn(3, 1, 64000, generator=g)
    st, ist = se.ConvSTFT(400, 100, 512, "hann", "complex"), se.ConviSTFT(400, 100, 512, 64000, "hann", "complex")
    specs = st(x.cuda())
    a = [t.cuda().requires_grad_(True) for t in (mre, mim)]
    y = ist.forward_masked(specs, a[0], a[1], mode)
    ga = torch.autograd.grad(y, a, gy.cuda())
    b = [t.cuda().requires_grad_(True) for t in (mre, mim)]
    y2 = ist(se.apply_mask_dccrn(specs, b[0], b[1], mode))
    gb = torch.autograd.grad(y2, b, gy.cuda())
    assert y.shape == y2.shape == (3, 1, 64000)
    assert rel(y, y2) < 1e-6
    for u, v in zip(ga, gb):
        assert rel(u, v) < 1e-5
    # oracle: the reference's mask expressions + conv_transpose1d ConviSTFT in float64
    c = [t.double().requires_grad_(True) for t in (mre, mim)]
    sp = specs.detach().cpu().double()
    o = oref.mask_apply_ref(torch.stack([sp[:, :257], sp[:, 257:]], -1), torch.stack([c[0], c[1]], -1), mode, False)
    y3 = oref.conv_istft_ref(torch.cat([o[..., 0], o[..., 1]], 1), 400, 100, 512, length=64000)
    gc = torch.autograd.grad(y3, c, gy.double().reshape(y3.shape))
    assert rel(y, y3.reshape(y.shape)) < TOL_SPEC
    # DC row excluded: the masks are zero-padded there (dccrn.py:200-201; F.pad drops that gradient) and the
    # reference's polar expression has a NaN derivative at an exactly-zero mask
    for u, v in zip(ga, gc):
        assert rel(u[:, 1:], v[:, 1:]) < TOL_GRAD
    # a spectrum that requires grad takes the two-stage path
    s2 = specs.detach().clone().requires_grad_(True)
    y4 = ist.forward_masked(s2, a[0], a[1], mode)
    (gs,) = torch.autograd.grad(y4, s2, gy.cuda())
    assert gs.shape == specs.shape and rel(y4, y2) < 1e-6


@pytest.mark.parametrize("shape", [(2, 1, 6000), (3, 2, 1, 16000)])
def test_mrstft_loss_and_gradient(se, oref, shape):
    g = torch.Generator().manual_seed(1236)
    ref = torch.randn(*shape, generator=g)
    est = ref + 0.1 * torch.randn(*shape, generator=g)
    e_ref = est.clone().requires_grad_(True)
    l_ref = oref.mrstft_loss_ref(e_ref, ref)
    (g_ref,) = torch.autograd.grad(l_ref, e_ref)
    e = est.cuda().requires_grad_(True)
    loss = se.loss_mrstft(e, ref.cuda())
    assert loss.dim() == 0
    (grad,) = torch.autograd.grad(2.0 * loss, e)
    assert abs(float(loss) - float(l_ref)) / float(l_ref) < TOL_GRAD
    # Ground truth is the float64 restatement.  north_star asks for gradients within 1e-3 relative;
    # on this loss the reference's OWN fp32 path (torch.stft + autograd) sits 0.5e-3 .. 2e-3 from
    # float64, because d/dA of log|A| is A/|A|^2 and amplifies fp32 FFT round-off in the few
    # near-silent bins.  So the bar is: within 1e-3 of float64, or at least as close to float64 as
    # the reference's fp32 path is (x1.25 for run-to-run spread), in max-norm and in relative L2.
    from oracle import spectral_np64 as o64
    l64, g64 = o64.mrstft_loss(est.numpy(), ref.numpy(), with_grad=True)
    g64 = torch.from_numpy(2.0 * g64).reshape(shape)
    assert abs(float(loss) - l64) / l64 < 1e-5

    err_ours, err_ref32 = rel(grad, g64), rel(2.0 * g_ref, g64)
    assert err_ours < max(TOL_GRAD, 2.0 * err_ref32), (err_ours, err_ref32)
    l2_ours, l2_ref32 = rel_l2(grad, g64), rel_l2(2.0 * g_ref, g64)
    print(f"mrstft grad {shape}: max-rel ours {err_ours:.2e} / reference fp32 {err_ref32:.2e}; rel-L2 ours {l2_ours:.2e} / reference fp32 {l2_ref32:.2e}")
    # (relative L2 is asserted at full size, test_cfg3_full_size_loss_and_gradient_vs_oracle; on these few-frame inputs it
    # swings between 0.3x and 3x of the reference fp32 path's own deviation from draw to draw)
    # what training consumes: directional derivatives against float64 -- along the gradient itself
    # (norm and direction, 1e-3 relative) and along random directions (error projects as
    # ||e|| ||v|| / sqrt(n); 6-sigma bound with ||e|| <= 2e-3 ||g||)
    directional_check(grad, g64, 2.0 * g_ref)


_MODE_SCRIPT = """
import sys, numpy as np, torch
sys.path.insert(0, {root!r})
import speech_enhancement_pytorch_b200 as se
g = torch.Generator().manual_seed(4242)
ref = torch.randn(5, 1, 20000, generator=g)
est = (ref + 0.2 * torch.randn(5, 1, 20000, generator=g)).cuda().requires_grad_(True)
loss = se.loss_mrstft(est, ref.cuda())
(grad,) = torch.autograd.grad(loss, est)
np.savez({out!r}, loss=float(loss), grad=grad.cpu().numpy())
"""


@pytest.mark.parametrize("env", [{"SE_MRSTFT_SAVE_SPECTRUM": "1"}, {"SE_MRSTFT_SAVE_SPECTRUM": "1", "SE_FORCE_GROUPS": "3"},
                                 {"SE_FORCE_GROUPS": "3"}])
def test_mrstft_loss_modes_agree(se, tmp_path, env):
    """The recompute backward (default), the opt-in saved-spectrum mode and forced multi-group chunking are the same
    function: each runs in its own process (the switches are read once) and must reproduce the default."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "mode.npz")
    subprocess.run([sys.executable, "-c", _MODE_SCRIPT.format(root=root, out=out)], check=True, env={**os.environ, **env},
                   timeout=600)
    got = np.load(out)
    g = torch.Generator().manual_seed(4242)
    ref = torch.randn(5, 1, 20000, generator=g)
    est = (ref + 0.2 * torch.randn(5, 1, 20000, generator=g)).cuda().requires_grad_(True)
    loss = se.loss_mrstft(est, ref.cuda())
    (grad,) = torch.autograd.grad(loss, est)
    assert abs(float(got["loss"]) - float(loss)) <= 1e-6 * abs(float(loss))
    assert rel(torch.from_numpy(got["grad"]), grad) < 2e-5



def test_mrstft_silent_and_identical_inputs(se, oref):
    """log / clamp corners (SURVEY 8c fixtures): all-zero signals sit on the 1e-7 clamp, identical signals give a zero
    numerator in the spectral-convergence term -- loss 0, gradient exactly 0, nothing non-finite; one silent row inside a
    normal batch matches the oracle."""
    z = torch.zeros(2, 1, 6000, device="cuda")
    e = z.clone().requires_grad_(True)
    loss = se.loss_mrstft(e, z)
    (g,) = torch.autograd.grad(loss, e)
    assert float(loss) == 0.0 and torch.count_nonzero(g) == 0
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 1, 6000, generator=gen)
    e = x.cuda().requires_grad_(True)
    loss = se.loss_mrstft(e, x.cuda())
    (g,) = torch.autograd.grad(loss, e)
    assert float(loss) == 0.0 and torch.isfinite(g).all() and float(g.abs().max()) == 0.0
    ref = torch.randn(3, 1, 6000, generator=gen)
    est = ref + 0.1 * torch.randn(3, 1, 6000, generator=gen)
    ref[1] = 0.0
    est[1] = 0.0
    e = est.cuda().requires_grad_(True)
    loss = se.loss_mrstft(e, ref.cuda())
    (g,) = torch.autograd.grad(loss, e)
    er = est.clone().double().requires_grad_(True)
    lr = oref.mrstft_loss_ref(er, ref.double())
    (gr,) = torch.autograd.grad(lr, er)
    assert torch.isfinite(g).all() and abs(float(loss) - float(lr)) < TOL_GRAD * float(lr)
    assert float(g[1].abs().max()) == 0.0 and float(gr[1].abs().max()) == 0.0
    assert rel(g, gr) < 2 * TOL_GRAD


def test_mrstft_full_size_survey_value(se):
    """seed 1236, est = ref + 0.1 N(0,1), 128x1x64000 -> loss 0.168027."""
    g = torch.Generator().manual_seed(1236)
    ref = torch.randn(128, 1, 64000, generator=g)
    est = ref + 0.1 * torch.randn(128, 1, 64000, generator=g)
    loss = se.loss_mrstft(est.cuda(), ref.cuda())
    assert abs(float(loss) - 0.168027) < 2e-4
    # determinism of the two-stage reduction
    assert float(se.loss_mrstft(est.cuda(), ref.cuda())) == float(loss)


def test_chain_matches_oracle_end_to_end(se, oref):
    """stft -> 'E' mask (pre-tanh) -> istft -> MR-STFT loss, gradient wrt the raw mask (cfg2 shape, small batch)."""
    g = torch.Generator().manual_seed(9)
    c = cfg(1024, 256, 1024)
    x = torch.randn(2, 1, 16384, generator=g)
    clean = x + 0.3 * torch.randn(2, 1, 16384, generator=g)
    raw = torch.randn(2, 1, 513, 65, 2, generator=g)

    def run(stft, mask, istft, loss, dev):
        r = raw.to(dev).requires_grad_(True)
        y = istft(mask(stft(x.to(dev), c), r), 16384, c)
        l = loss(y, clean.to(dev))
        return y, l, torch.autograd.grad(l, r)[0]

    y, l, gr = run(se.stft_custom, lambda s, m: se.apply_mask(s, m, "E", True), se.istft_custom, se.loss_mrstft, "cuda")
    y0, l0, gr0 = run(oref.stft_custom_ref, lambda s, m: oref.mask_apply_ref(s, m, "E", True), oref.istft_custom_ref,
                      oref.mrstft_loss_ref, "cpu")
    assert rel(y, y0) < TOL_SPEC
    assert abs(float(l) - float(l0)) / float(l0) < TOL_GRAD
    # gradient: float64 ground truth through the same chain (torch CPU float64), see the note above
    x64, c64, r64 = x.double(), clean.double(), raw.double().requires_grad_(True)
    y64 = oref.istft_custom_ref(oref.mask_apply_ref(oref.stft_custom_ref(x64, c), r64, "E", True), 16384, c)
    (g64,) = torch.autograd.grad(oref.mrstft_loss_ref(y64, c64), r64)
    err_ours, err_ref32 = rel(gr, g64), rel(gr0, g64)
    assert err_ours < max(TOL_GRAD, 2.0 * err_ref32), (err_ours, err_ref32)      # max-norm is ill-conditioned here
    directional_check(gr, g64, gr0)


def test_errors_match_reference_behaviour(se):
    x = torch.randn(1, 1, 4096, device="cuda")
    with pytest.raises(NotImplementedError):
        se.stft_custom(x, cfg(321, 80, 321))                                         # odd n_fft (even sizes run, see test_gpu_generic.py)
    with pytest.raises(TypeError):
        se.stft_custom(x.int(), cfg(512, 128, 512))                                  # float64 is accepted now (see the fp64 test)
    with pytest.raises((ValueError, RuntimeError)):
        se.stft_custom(torch.randn(1, 1, 200, device="cuda"), cfg(512, 128, 512))   # reflect pad >= N
    spec = se.stft_custom(x.bfloat16(), cfg(512, 128, 512))                          # bf16 in -> fp32 spectra
    assert spec.dtype == torch.float32


@pytest.mark.parametrize("name", ["conv_a", "conv_b", "conv_c"])
def test_conv_stft_golden(se, name):
    g = golden(name)
    wl, inc, nfft, length = (int(v) for v in g["meta"])
    st = se.ConvSTFT(wl, inc, nfft, "hann", "complex")
    spec = st(torch.from_numpy(g["x"]).cuda())
    assert spec.shape == g["spec"].shape
    assert rel(spec, torch.from_numpy(g["spec"])) < TOL_SPEC


@pytest.mark.parametrize("name", ["conv_a", "conv_b", "conv_c"])
def test_conv_istft_golden_and_autograd(se, oref, name):
    g = golden(name)
    wl, inc, nfft, length = (int(v) for v in g["meta"])
    length = None if length < 0 else length
    ist = se.ConviSTFT(wl, inc, nfft, length, "hann", "complex")
    for ks, ky in (("spec", "y"), ("spec2", "y2")):
        y = ist(torch.from_numpy(g[ks]).cuda())
        assert y.shape == g[ky].shape
        assert rel(y, torch.from_numpy(g[ky])) < TOL_SPEC
    s = torch.from_numpy(g["spec2"]).cuda().requires_grad_(True)
    y = ist(s)
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(1))
    (gs,) = torch.autograd.grad(y, s, gy.cuda())
    sr = torch.from_numpy(g["spec2"]).double().requires_grad_(True)
    (gsr,) = torch.autograd.grad(oref.conv_istft_ref(sr, wl, inc, nfft, "hann", length), sr, gy.double())
    assert rel(gs, gsr) < TOL_SPEC


def test_conv_polar_golden(se):
    g = golden("conv_polar")
    st = se.ConvSTFT(400, 100, 512, "hann", "real")
    mags, phase = st(torch.from_numpy(g["x"]).cuda())
    assert rel(mags, torch.from_numpy(g["mags"])) < TOL_SPEC
    ist = se.ConviSTFT(400, 100, 512, None, "hann", "real")
    y = ist(torch.from_numpy(g["mags"]).cuda(), torch.from_numpy(g["phase"]).cuda())
    assert rel(y, torch.from_numpy(g["y"])) < TOL_SPEC


@pytest.mark.parametrize("win_type", ["hamming", "blackman", None])
def test_dccrn_modules_with_any_scipy_window(se, oref, win_type):
    """The reference's constructors default to win_type='hamming' and take any scipy window (src/model/dccrn.py:651-655,
    671,705): analysis, synthesis, autograd and the fused mask tail against the oracle's dense conv path."""
    g = torch.Generator().manual_seed(31)
    x = torch.randn(3, 1, 6000, generator=g)
    st = se.ConvSTFT(400, 100, 512, win_type, "complex").cuda()
    ist = se.ConviSTFT(400, 100, 512, 6000, win_type, "complex").cuda()
    spec = st(x.cuda())
    want = oref.conv_stft_ref(x, 400, 100, 512, win_type)
    assert rel(spec, want) < TOL_SPEC
    s = want.clone().requires_grad_(True)
    yw = oref.conv_istft_ref(s, 400, 100, 512, win_type, length=6000)
    gy = torch.randn(yw.shape, generator=g)
    (gw,) = torch.autograd.grad(yw, s, gy)
    s2 = want.cuda().requires_grad_(True)
    y = ist(s2)
    (gs,) = torch.autograd.grad(y, s2, gy.cuda())
    assert rel(y, yw) < TOL_SPEC and rel(gs, gw) < TOL_GRAD
    mre, mim = torch.randn(3, 257, spec.shape[-1], generator=g), torch.randn(3, 257, spec.shape[-1], generator=g)
    masked = oref.mask_apply_ref(torch.stack([want[:, :257], want[:, 257:]], -1), torch.stack([mre, mim], -1), "C")
    yw2 = oref.conv_istft_ref(torch.cat([masked[..., 0], masked[..., 1]], 1), 400, 100, 512, win_type, length=6000)
    assert rel(ist.forward_masked(spec, mre.cuda(), mim.cuda(), "C"), yw2) < TOL_SPEC
    # a reference state dict (same buffer names and shapes) loads strictly
    ist.load_state_dict({k: v.clone() for k, v in ist.state_dict().items()}, strict=True)


def test_dccrn_real_feature_path_is_fused_and_differentiable(se, oref):
    """feature_type='real': ConvSTFT returns (mags, phase) (dccrn.py:696-701) and ConviSTFT takes (mags, phase) (:729-732);
    one launch each, gradient to both."""
    g = torch.Generator().manual_seed(32)
    x = torch.randn(2, 1, 5000, generator=g)
    st = se.ConvSTFT(400, 100, 512, "hann", "real").cuda()
    ist = se.ConviSTFT(400, 100, 512, 5000, "hann", "real").cuda()
    mags, phase = st(x.cuda())
    wm, wp = oref.conv_stft_ref(x, 400, 100, 512, "hann", "real")
    assert rel(mags, wm) < TOL_SPEC
    assert float(torch.remainder(phase.cpu() - wp + math.pi, 2 * math.pi).sub(math.pi).abs().mul(wm > 1e-3).max()) < 1e-4
    m1, p1 = wm.clone().requires_grad_(True), wp.clone().requires_grad_(True)
    yw = oref.conv_istft_ref(m1, 400, 100, 512, "hann", length=5000, phase=p1)
    gy = torch.randn(yw.shape, generator=g)
    gmw, gpw = torch.autograd.grad(yw, (m1, p1), gy)
    m2, p2 = wm.cuda().requires_grad_(True), wp.cuda().requires_grad_(True)
    y = ist(m2, p2)
    gm, gp = torch.autograd.grad(y, (m2, p2), gy.cuda())
    assert rel(y, yw) < TOL_SPEC and rel(gm, gmw) < TOL_GRAD and rel(gp, gpw) < TOL_GRAD


def test_float64_inputs_are_accepted_like_the_reference(se, oref):
    """The reference passes dtype=tensor.dtype (src/evaluate.py:113,147) and works in float64; here float64 inputs are
    computed in fp32 and returned as float64: tolerance 1e-5 relative against the float64 oracle."""
    c = cfg(512, 128, 512)
    g = torch.Generator().manual_seed(33)
    x = torch.randn(2, 1, 8000, generator=g, dtype=torch.float64)
    spec = se.stft_custom(x.cuda(), c)
    assert spec.dtype == torch.float64
    want = oref.stft_custom_ref(x, c)
    assert rel(spec, want) < 1e-5
    y = se.istft_custom(spec, 8000, c)
    assert y.dtype == torch.float64 and rel(y, x) < 1e-5
    xg = x.cuda().requires_grad_(True)
    (gx,) = torch.autograd.grad(se.stft_custom(xg, c).square().sum(), xg)
    xr = x.clone().requires_grad_(True)
    (gr,) = torch.autograd.grad(oref.stft_custom_ref(xr, c).square().sum(), xr)
    assert gx.dtype == torch.float64 and rel(gx, gr) < 1e-5


@pytest.mark.parametrize("mode", ["E", "C", "R"])
def test_dccrn_wave_to_wave_golden(se, mode):
    """Real DCCRN forward captured by hooks: specs = stft(x); wav = clamp(istft(mask(specs)))."""
    g = golden(f"dccrn_mask_{mode}")
    st = se.ConvSTFT(400, 100, 512, "hann", "complex")
    ist = se.ConviSTFT(400, 100, 512, 1600, "hann", "complex")
    specs = st(torch.from_numpy(g["x"]).cuda())
    assert rel(specs, torch.from_numpy(g["specs"])) < TOL_SPEC
    out_spec = se.apply_mask_dccrn(specs, torch.from_numpy(g["mask_re"]).cuda(), torch.from_numpy(g["mask_im"]).cuda(), mode)
    wav = torch.clamp_(ist(out_spec), -1, 1)          # dccrn.py:228 clamps the module output in place
    assert wav.shape == g["wav"].shape
    assert rel(wav, torch.from_numpy(g["wav"])) < TOL_SPEC


@pytest.mark.parametrize("n,h,w,N", [(512, 128, 512, 16000), (1024, 256, 1024, 16384), (2048, 512, 2048, 44100),
                                     (512, 128, 400, 9999)])
@pytest.mark.parametrize("mode", ["real", "E", "C", "R"])
def test_fused_enhance_matches_unfused_and_oracle(se, oref, n, h, w, N, mode):
    g = torch.Generator().manual_seed(n + N)
    c = cfg(n, h, w)
    pre_tanh = mode == "E"
    x = torch.randn(3, 1, N, generator=g)
    F, T = n // 2 + 1, 1 + N // h
    mask = torch.randn(*((3, 1, F, T) if mode == "real" else (3, 1, F, T, 2)), generator=g)
    mr = mask.clone().requires_grad_(True)
    ref = oref.istft_custom_ref(oref.mask_apply_ref(oref.stft_custom_ref(x, c), mr, mode, pre_tanh), N, c)
    mc = mask.cuda().requires_grad_(True)
    y = se.enhance(x.cuda(), mc, c, mode, pre_tanh)
    assert rel(y, ref) < TOL_SPEC
    gy = torch.randn(ref.shape, generator=g)
    (gref,) = torch.autograd.grad(ref, mr, gy)
    (gm,) = torch.autograd.grad(y, mc, gy.cuda())
    assert rel(gm, gref) < TOL_GRAD
    m2 = mask.cuda().requires_grad_(True)
    y2 = se.istft_custom(se.apply_mask(se.stft_custom(x.cuda(), c), m2, mode, pre_tanh), N, c)
    assert rel(y, y2) < 1e-5
    (gm2,) = torch.autograd.grad(y2, m2, gy.cuda())
    assert rel(gm2, gref) < TOL_GRAD


@pytest.mark.parametrize("n,h,w,N,length", [(512, 128, 512, 16000, 16000), (512, 256, 512, 12345, 12345),
                                            (1024, 256, 1024, 16384, 16000), (1024, 512, 1024, 30000, 30000),
                                            (2048, 512, 2048, 44100, 44100), (2048, 1024, 2048, 40000, 39000),
                                            (512, 128, 400, 9999, 9999)])
@pytest.mark.parametrize("mode", ["real", "E", "C", "R"])
def test_mask_istft_tail_matches_two_stage_and_oracle(se, oref, n, h, w, N, length, mode):
    """apply_mask_istft == istft_custom(apply_mask(.)) (the masked spectrum is never written), fwd and bwd."""
    g = torch.Generator().manual_seed(n + N + len(mode))
    c = cfg(n, h, w)
    pre_tanh = mode in ("E", "R")
    F, T = n // 2 + 1, 1 + N // h
    spec = torch.randn(2, 2, F, T, 2, generator=g)
    mask = torch.randn(*((2, 2, F, T) if mode == "real" else (2, 2, F, T, 2)), generator=g)
    mr = mask.double().requires_grad_(True)
    ref = oref.istft_custom_ref(oref.mask_apply_ref(spec.double(), mr, mode, pre_tanh), length, c)
    mc = mask.cuda().requires_grad_(True)
    y = se.apply_mask_istft(spec.cuda(), mc, length, c, mode, pre_tanh)
    assert y.shape == ref.shape
    assert rel(y, ref) < TOL_SPEC
    gy = torch.randn(ref.shape, generator=g)
    (gref,) = torch.autograd.grad(ref, mr, gy.double())
    (gm,) = torch.autograd.grad(y, mc, gy.cuda())
    assert rel(gm, gref) < TOL_SPEC
    m2 = mask.cuda().requires_grad_(True)
    y2 = se.istft_custom(se.apply_mask(spec.cuda(), m2, mode, pre_tanh), length, c)
    assert rel(y, y2) < 1e-5
    (gm2,) = torch.autograd.grad(y2, m2, gy.cuda())
    assert rel(gm, gm2) < 1e-5


def test_mask_istft_tail_spectrum_requiring_grad_falls_back_to_two_stages(se):
    c = cfg(512, 128, 512)
    spec = torch.randn(1, 1, 257, 40, 2, device="cuda", requires_grad=True)
    mask = torch.randn(1, 1, 257, 40, 2, device="cuda", requires_grad=True)
    y = se.apply_mask_istft(spec, mask, 4992, c, "C")
    gs, gm = torch.autograd.grad(y.square().sum(), (spec, mask))
    s2, m2 = spec.detach().requires_grad_(True), mask.detach().requires_grad_(True)
    y2 = se.istft_custom(se.apply_mask(s2, m2, "C"), 4992, c)
    gs2, gm2 = torch.autograd.grad(y2.square().sum(), (s2, m2))
    assert torch.equal(y, y2) and torch.equal(gs, gs2) and torch.equal(gm, gm2)


def test_mask_istft_tail_full_size_cfg2(se):
    """cfg2 at full size: the fused tail and the two-stage path agree, forward and backward."""
    c = cfg(1024, 256, 1024)
    spec = torch.randn(64, 1, 513, 251, 2, device="cuda")
    raw = torch.randn(64, 1, 513, 251, 2, device="cuda")
    gy = torch.randn(64, 1, 64000, device="cuda")
    r1, r2 = raw.clone().requires_grad_(True), raw.clone().requires_grad_(True)
    y1 = se.apply_mask_istft(spec, r1, 64000, c, "E", True)
    y2 = se.istft_custom(se.apply_mask(spec, r2, "E", True), 64000, c)
    assert float((y1 - y2).abs().max()) < 1e-5 * float(y2.abs().max())
    (g1,) = torch.autograd.grad(y1, r1, gy)
    (g2,) = torch.autograd.grad(y2, r2, gy)
    assert float((g1 - g2).abs().max()) < 1e-5 * float(g2.abs().max())


def test_fused_chain_full_size_cfg2(se):
    """cfg2 at full size: fused and unfused paths agree (size-independent consistency)."""
    c = cfg(1024, 256, 1024)
    x = torch.randn(64, 1, 64000, device="cuda")
    raw = torch.randn(64, 1, 513, 251, 2, device="cuda")
    y1 = se.enhance(x, raw, c, "E", True)
    y2 = se.istft_custom(se.apply_mask(se.stft_custom(x, c), raw, "E", True), 64000, c)
    assert float((y1 - y2).abs().max()) < 1e-5 * float(y2.abs().max())


@pytest.mark.parametrize("name,model", [("dnn", None), ("dnn", "mask"), ("demucs", None)])
def test_evaluate_matches_reference_flow(se, oref, name, model):
    """evaluate(): z-score, segment (stride = win_length), STFT, model, iSTFT, stitch (src/evaluate.py:10-98)."""
    g = torch.Generator().manual_seed(21)
    mix = torch.randn(1, 2, 30000, generator=g) * 0.3 + 0.05
    config = types.SimpleNamespace(
        dset=types.SimpleNamespace(norm="z-score", sample_rate=16000),
        model=types.SimpleNamespace(name=name, segment=0.512, n_fft=512, hop_length=128, win_length=512, center=True))
    fn = None
    if model:
        class Toy(torch.nn.Module):                      # a spectrogram "model": fixed smooth real mask
            def forward(self, spec):
                f = torch.linspace(0.2, 1.0, spec.shape[-3], device=spec.device)[:, None, None]
                return spec * f
        fn = Toy()
    want = oref.evaluate_ref(mix, fn, config)
    got = se.evaluate(mix, fn.cuda() if fn else None, "cuda", config)
    assert got.shape == want.shape
    assert rel(got, want) < TOL_SPEC


def test_evaluate_identity_property_like_reference_test(se):
    """test/test_eval.py:5-40: evaluate with no model is the identity (error_rate 1e-10 there, fp32 here)."""
    mix = torch.randn(1, 1, 16000 * 3, device="cuda")
    config = types.SimpleNamespace(
        dset=types.SimpleNamespace(norm="z-score", sample_rate=16000),
        model=types.SimpleNamespace(name="dnn", segment=1.024, n_fft=512, hop_length=128, win_length=512, center=True))
    out = se.evaluate(mix, None, "cuda", config)
    assert float((out - mix).abs().max()) < 1e-5


def test_cfg5_long_form_44k_stereo(se, oref):
    """BASELINE cfg 5 shape: 44.1 kHz stereo 30 s clips (N = 1 323 000), n_fft 2048 / 1024, complex mask."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 2, 1323000, generator=g)
    for n in (2048, 1024):
        c = cfg(n, n // 4, n)
        xs = x.cuda()
        spec = se.stft_custom(xs, c)
        assert spec.shape == (1, 2, n // 2 + 1, 1 + 1323000 // (n // 4), 2)
        y = se.istft_custom(spec, 1323000, c)
        assert float((y - xs).abs().max()) < 2e-5
        # spot-check the spectrum against the oracle on a slice of frames (the oracle on 30 s is slow but fine)
        ref = oref.stft_custom_ref(x[:, :1, :200000], c)
        got = se.stft_custom(xs[:, :1, :200000].contiguous(), c)
        assert rel(got, ref) < TOL_SPEC
        mask = torch.rand(1, 2, n // 2 + 1, spec.shape[-2], 2, generator=g) * 2 - 1
        y1 = se.enhance(xs, mask.cuda(), c, "C")
        y2 = se.istft_custom(se.apply_mask(spec, mask.cuda(), "C"), 1323000, c)
        assert float((y1 - y2).abs().max()) < 1e-5 * max(1.0, float(y2.abs().max()))


def test_cfg4_dccrn_transforms_full_size(se):
    """BASELINE cfg 4 shape: 16 x 4 s through ConvSTFT -> ConviSTFT(length=64000) ~ identity (SURVEY: 1.7e-6)."""
    x = torch.randn(16, 1, 64000, device="cuda")
    st = se.ConvSTFT(400, 100, 512, "hann", "complex")
    ist = se.ConviSTFT(400, 100, 512, 64000, "hann", "complex")
    spec = st(x)
    assert spec.shape == (16, 514, 643)
    y = ist(spec)
    assert y.shape == (16, 1, 64000)
    assert float((y - x).abs().max()) < 2e-5
    with torch.autocast("cuda", dtype=torch.bfloat16):          # bf16 model, fp32 spectra
        y16 = ist(st(x.bfloat16()))
    assert y16.dtype == torch.float32


def test_cfg1_and_cfg3_shapes(se, oref):
    """cfg 1 (16x4 s, 512/128, real mask) and cfg 3 (128x4 s MR-STFT loss fwd+bwd) run at full size."""
    c = cfg(512, 128, 512)
    x = torch.randn(16, 1, 64000, device="cuda")
    m = torch.rand(16, 1, 257, 501, device="cuda")
    y = se.istft_custom(se.apply_mask(se.stft_custom(x, c), m, "real"), 64000, c)
    y2 = se.enhance(x, m, c, "real")
    assert float((y - y2).abs().max()) < 1e-5
    ref = oref.istft_custom_ref(oref.mask_apply_ref(oref.stft_custom_ref(x[:2].cpu(), c), m[:2].cpu(), "real"), 64000, c)
    assert rel(y[:2], ref) < TOL_SPEC
    est = torch.randn(128, 1, 64000, device="cuda", requires_grad=True)
    tgt = torch.randn(128, 1, 64000, device="cuda")
    loss = se.loss_mrstft(est, tgt)
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(est.grad).all()
    # linearity of the backward in the upstream gradient
    est2 = est.detach().clone().requires_grad_(True)
    (3.0 * se.loss_mrstft(est2, tgt)).backward()
    assert rel(est2.grad, 3.0 * est.grad) < 1e-5


def test_cfg3_full_size_loss_and_gradient_vs_oracle(se, oref):
    """BASELINE cfg3 at FULL size (128 x 4 s, SURVEY 8d inputs: seed 1236, est = ref + 0.1 N(0,1)): loss and gradient of the
    CUDA path against the oracle evaluated in float64 on the CPU (the exact value of the reference-convention loss), with
    the reference's own fp32 torch path measured beside it.  north_star: loss and gradients within 1e-3 relative."""
    g = torch.Generator().manual_seed(1236)
    ref = torch.randn(128, 1, 64000, generator=g)
    est = ref + 0.1 * torch.randn(128, 1, 64000, generator=g)
    e = est.cuda().requires_grad_(True)
    loss = se.loss_mrstft(e, ref.cuda())
    (grad,) = torch.autograd.grad(loss, e)
    e64 = est.double().requires_grad_(True)
    l64 = oref.mrstft_loss_ref(e64, ref.double())
    (g64,) = torch.autograd.grad(l64, e64)
    e32 = est.clone().requires_grad_(True)
    l32 = oref.mrstft_loss_ref(e32, ref)
    (g32,) = torch.autograd.grad(l32, e32)
    got = {"loss": float(loss), "loss_f64": float(l64), "loss_rel_err": abs(float(loss) - float(l64)) / float(l64),
           "grad_rel_l2": rel_l2(grad, g64), "grad_max_rel": rel(grad, g64),
           "reference_fp32_grad_rel_l2": rel_l2(g32, g64), "reference_fp32_grad_max_rel": rel(g32, g64),
           "grad_rel_l2_vs_reference_fp32": rel_l2(grad, g32)}
    print("cfg3 full size:", got)
    assert abs(float(l64) - 0.168027) < 2e-4                     # SURVEY section 6's value for these inputs
    assert got["loss_rel_err"] < 1e-6
    # The bar for the gradient: 1e-3 of float64 where fp32 can deliver it, else no further from float64 than the
    # reference's own fp32 torch path is (x1.25): d log|A| / dA = A/|A|^2 amplifies fp32 FFT round-off in near-silent bins,
    # and at this size BOTH fp32 paths sit ~2e-3 from float64 (measured: ours 2.2e-3, torch.stft + autograd 2.1e-3).
    assert got["grad_rel_l2"] < max(TOL_GRAD, 1.25 * got["reference_fp32_grad_rel_l2"]), got
    assert got["grad_max_rel"] < max(TOL_GRAD, 1.25 * got["reference_fp32_grad_max_rel"]), got


def test_cfg2_full_size_chain_vs_oracle(se, oref):
    """BASELINE cfg2 at FULL size (64 x 4 s, n_fft 1024 / hop 256): stft_custom -> DCUnet 'E' mask (tanh) -> istft_custom ->
    MR-STFT loss -> gradient to the raw mask, against the oracle chain on the CPU (float64 ground truth, the reference's
    fp32 path beside it)."""
    c = cfg(1024, 256, 1024)
    g = torch.Generator().manual_seed(1235)
    x = torch.randn(64, 1, 64000, generator=g)
    clean = x + 0.3 * torch.randn(64, 1, 64000, generator=g)
    raw = torch.randn(64, 1, 513, 251, 2, generator=g)

    def chain(stft, mask, istft, loss_fn, x, clean, raw):
        raw = raw.clone().requires_grad_(True)
        y = istft(mask(stft(x, c), raw), 64000, c)
        l = loss_fn(y, clean)
        (graw,) = torch.autograd.grad(l, raw)
        return y.detach(), l.detach(), graw

    ours = chain(se.stft_custom, lambda s, m: se.apply_mask(s, m, "E", True), se.istft_custom, se.loss_mrstft,
                 x.cuda(), clean.cuda(), raw.cuda())
    fused = chain(se.stft_custom, lambda s, m: (s, m), lambda sm, n, cc: se.apply_mask_istft(sm[0], sm[1], n, cc, "E", True),
                  se.loss_mrstft, x.cuda(), clean.cuda(), raw.cuda())
    ref_fn = (oref.stft_custom_ref, lambda s, m: oref.mask_apply_ref(s, m, "E", True), oref.istft_custom_ref, oref.mrstft_loss_ref)
    r64 = chain(*ref_fn, x.double(), clean.double(), raw.double())
    r32 = chain(*ref_fn, x, clean, raw)
    for name, got in (("dropin", ours), ("fused tail", fused)):
        res = {"wave_max_rel": rel(got[0], r64[0]), "loss_rel_err": abs(float(got[1]) - float(r64[1])) / float(r64[1]),
               "grad_rel_l2": rel_l2(got[2], r64[2]), "grad_max_rel": rel(got[2], r64[2]),
               "reference_fp32_grad_rel_l2": rel_l2(r32[2], r64[2]), "reference_fp32_grad_max_rel": rel(r32[2], r64[2])}
        print(f"cfg2 full size ({name}):", res)
        assert res["wave_max_rel"] < TOL_SPEC, res
        assert res["loss_rel_err"] < TOL_GRAD, res
        assert res["grad_rel_l2"] < max(TOL_GRAD, 1.25 * res["reference_fp32_grad_rel_l2"]), res
        assert res["grad_max_rel"] < max(TOL_GRAD, 1.25 * res["reference_fp32_grad_max_rel"]), res


def test_f_rows_match_reference_run_goldens(se):
    """SURVEY 8f rows against fixtures produced by the REAL reference (tests/golden/make_golden.py): src.loss.si_snr /
    loss_sisdr / PSA, src.evaluate.evaluate(model=None), the models' magnitude features."""
    gd = golden("losses")
    for tag in "abc":
        s1, s2 = torch.from_numpy(gd[f"sisnr_s1_{tag}"]).cuda(), torch.from_numpy(gd[f"sisnr_s2_{tag}"]).cuda()
        assert abs(float(se.si_snr(s1, s2)) - float(gd[f"sisnr_{tag}"])) < 1e-5 * abs(float(gd[f"sisnr_{tag}"]))
        assert abs(float(se.loss_sisdr(s1, s2)) - float(gd[f"sisdr_loss_{tag}"])) < 1e-5 * abs(float(gd[f"sisdr_loss_{tag}"]))
    enh, tgt, mix = (torch.from_numpy(gd[k]).cuda() for k in ("psa_enh", "psa_tgt", "psa_mix"))
    assert abs(float(se.loss_phase_sensitive_spectral_approximation(enh, tgt, mix)) - float(gd["psa"])) < 1e-5 * float(gd["psa"])
    gd = golden("features")
    spec = torch.from_numpy(gd["spec"]).cuda()
    for kind in ("amplitude", "power", "magnitude"):
        assert rel(se.magnitude_feature(spec, kind), torch.from_numpy(gd[kind])) < 1e-6, kind
    got = se.magnitude_feature(torch.from_numpy(gd["crn_spec"]).cuda(), "crn").cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(gd["crn"]))
    assert np.allclose(np.nan_to_num(got), np.nan_to_num(gd["crn"]), rtol=1e-5, atol=1e-6)
    gd = golden("evaluate")
    for tag in ("zscore", "plain", "n1024"):
        n_fft, hop, nfeat, zscore = (int(v) for v in gd[f"meta_{tag}"])
        conf = types.SimpleNamespace(dset=types.SimpleNamespace(norm="z-score" if zscore else "none", sample_rate=16000),
                                     model=types.SimpleNamespace(name="unet", n_fft=n_fft, hop_length=hop, win_length=n_fft, center=True,
                                                                 segment=nfeat / 16000.0, sources=["clean"]))
        out = se.evaluate(torch.from_numpy(gd[f"mix_{tag}"]), None, "cuda", conf)
        assert tuple(out.shape) == gd[f"enh_{tag}"].shape
        assert rel(out, torch.from_numpy(gd[f"enh_{tag}"])) < 1e-5, tag


@pytest.mark.parametrize("kind", ["power", "magnitude", "amplitude", "crn"])
def test_magnitude_features(se, oref, kind):
    """NN input features, quirks kept (|re^2-im^2|, sqrt(re^2-im^2) with its NaNs)."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 1, 16000, generator=g)
    c = cfg(512, 128, 512)
    ref_spec = oref.stft_custom_ref(x, c)
    want = oref.magnitude_feature_ref(ref_spec, kind)
    spec, feat = se.stft_custom_with_feature(x.cuda(), c, kind)
    assert rel(spec, ref_spec) < TOL_SPEC
    got2 = se.magnitude_feature(ref_spec.cuda(), kind)
    for got in (feat, got2):
        assert got.shape == want.shape
        g_ = got.cpu()
        if kind == "crn":
            # sign of re^2-im^2 can flip within round-off: compare where the reference is comfortably real
            safe = (ref_spec[..., 0] ** 2 - ref_spec[..., 1] ** 2) > 1e-6 * (ref_spec ** 2).sum(-1)
            assert torch.isnan(got2.cpu()).eq(torch.isnan(want)).all()
            assert float((g_[safe] - want[safe]).abs().max() / want[safe].abs().max()) < 1e-3
        else:
            assert rel(g_, want) < TOL_SPEC


@pytest.mark.parametrize("kind,fn", [("mse", torch.nn.functional.mse_loss), ("l1", torch.nn.functional.l1_loss)])
@pytest.mark.parametrize("n,h", [(512, 128), (1024, 256)])
def test_fused_spectral_loss(se, oref, kind, fn, n, h):
    """8f-2: loss_function(enhanced, stft_custom(sources)) with torch's mse/l1 (src/distrib.py:263-267)."""
    g = torch.Generator().manual_seed(n)
    c = cfg(n, h, n)
    src = torch.randn(3, 1, 16000, generator=g)
    tspec = oref.stft_custom_ref(src, c)
    enh = tspec + 0.05 * torch.randn(tspec.shape, generator=g)
    er = enh.double().requires_grad_(True)
    want = fn(er, tspec.double())
    (gw,) = torch.autograd.grad(want, er)
    ec = enh.cuda().requires_grad_(True)
    got = se.loss_spectral(ec, src.cuda(), c, kind)
    assert got.dim() == 0
    assert abs(float(got) - float(want)) / float(want) < 1e-4
    (gg,) = torch.autograd.grad(got, ec)
    if kind == "mse":
        assert rel(gg, gw) < TOL_GRAD
    else:
        assert float((torch.sign(gg.cpu()) == torch.sign(gw)).double().mean()) > 0.999
        assert rel(gg.abs(), gw.abs()) < 1e-5


@pytest.mark.parametrize("shape,noise", [((4, 1, 16000), 0.3), ((2, 2, 1, 64000), 0.01), ((3, 1, 4001), 1.0)])
def test_si_snr_and_gradient(se, oref, shape, noise):
    """8f-4: si_snr / loss_sisdr (src/loss.py:14-29) in one pass, gradient A s1 + B s2."""
    g = torch.Generator().manual_seed(shape[-1])
    tgt = torch.randn(*shape, generator=g)
    est = tgt * 0.7 + noise * torch.randn(*shape, generator=g)
    er = est.double().requires_grad_(True)
    want = oref.si_snr_ref(er, tgt.double())
    (gw,) = torch.autograd.grad(-want, er)
    ec = est.cuda().requires_grad_(True)
    got = se.loss_sisdr(ec, tgt.cuda())
    assert got.dim() == 0
    assert abs(float(got) + float(want)) < 1e-3 * max(1.0, abs(float(want)))
    (gg,) = torch.autograd.grad(got, ec)
    assert rel(gg, gw) < TOL_GRAD
    # against the reference's own fp32 evaluation
    assert abs(float(got) + float(oref.si_snr_ref(est, tgt))) < 1e-3 * max(1.0, abs(float(want)))


@pytest.mark.parametrize("tag", ["half", "quarter", "coprime", "gap", "abut"])
def test_overlap_and_add_matches_reference_golden(se, oref, tag):
    """8f-4: Conv-TasNet's overlap_and_add (src/model/conv_tasnet.py:11-31): bit-exact, gradient = exact gather."""
    g = golden("tasnet_metric")
    sig, step = torch.from_numpy(g[f"sig_{tag}"]), int(g[f"step_{tag}"])
    sc = sig.cuda().requires_grad_(True)
    out = se.overlap_and_add(sc, step)
    assert out.shape == g[f"out_{tag}"].shape
    assert np.array_equal(out.detach().cpu().numpy(), g[f"out_{tag}"])
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    sr = sig.clone().requires_grad_(True)
    (gw,) = torch.autograd.grad(oref.overlap_and_add_ref(sr, step), sr, go)
    (gg,) = torch.autograd.grad(out, sc, go.cuda())
    assert torch.equal(gg.cpu(), gw)


def test_overlap_and_add_conv_tasnet_size(se, oref):
    """Decoder output of Conv-TasNet for a 4 s batch: [M, C, K, L] = [8, 2, 3199, 40], step L/2 (conv_tasnet.py:203)."""
    sig = torch.randn(8, 2, 3199, 40, generator=torch.Generator().manual_seed(5))
    out = se.overlap_and_add(sig.cuda(), 20)
    assert out.shape == (8, 2, 64000)
    assert torch.equal(out.cpu(), oref.overlap_and_add_ref(sig, 20))


def test_si_sdr_metric_matches_reference_golden(se, oref):
    """8f-4: SI_SDR (src/metric.py:92-123) from one pass over the two waveforms."""
    g = golden("tasnet_metric")
    ref, est = torch.from_numpy(g["sdr_ref"]).cuda(), torch.from_numpy(g["sdr_est"]).cuda()
    got = se.SI_SDR(ref, est)
    assert got.dim() == 0 and got.is_cuda
    assert abs(float(got) - float(g["sdr"])) < 1e-4
    assert abs(float(se.SI_SDR(ref, 0.01 * est + 0.5)) - float(g["sdr_scaled"])) < 1e-4
    big_r = torch.randn(4, 1, 64000, generator=torch.Generator().manual_seed(2))
    big_e = big_r * 0.5 + 0.2 * torch.randn(4, 1, 64000, generator=torch.Generator().manual_seed(3))
    assert abs(float(se.SI_SDR(big_r.cuda(), big_e.cuda())) - float(oref.si_sdr_metric_ref(big_r.double(), big_e.double()))) < 1e-4


def test_reentrant_from_threads_on_separate_streams(se, oref):
    """nn.DataParallel runs DCCRN's transforms from one Python thread per GPU (src/solver.py:145): the C-ABI
    must be re-entrant.  Two threads, two streams, different configurations, results checked per thread."""
    import threading
    g = torch.Generator().manual_seed(12)
    xs = [torch.randn(4, 1, 16000, generator=g) for _ in range(2)]
    cfgs = [cfg(512, 128, 512), cfg(1024, 256, 1024)]
    refs = [oref.stft_custom_ref(x, c) for x, c in zip(xs, cfgs)]
    out, err = [None, None], []

    def work(i):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                x = xs[i].cuda()
                for _ in range(20):
                    spec = se.stft_custom(x, cfgs[i])
                    y = se.istft_custom(spec, 16000, cfgs[i])
                stream.synchronize()
                out[i] = (spec.cpu(), float((y - x).abs().max()))
        except Exception as e:      # noqa: BLE001
            err.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not err, err
    for i in range(2):
        assert rel(out[i][0], refs[i]) < TOL_SPEC
        assert out[i][1] < 1e-5


def test_cuda_graph_capture_and_replay(se):
    """The launch path has no driver calls after warm-up (tables cached, smem opt-in done), so a whole
    forward chain -- including the programmatic-dependent-launch edges -- captures into a CUDA graph."""
    c = cfg(1024, 256, 1024)
    x = torch.randn(8, 1, 16000, device="cuda")
    raw = torch.randn(8, 1, 513, 63, 2, device="cuda")
    with torch.no_grad():
        ref = se.istft_custom(se.apply_mask(se.stft_custom(x, c), raw, "E", True), 16000, c)     # warm-up
        ref2 = se.enhance(x, raw, c, "E", True)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            y = se.istft_custom(se.apply_mask(se.stft_custom(x, c), raw, "E", True), 16000, c)
            y2 = se.enhance(x, raw, c, "E", True)
        x.copy_(torch.randn_like(x))
        graph.replay()
        torch.cuda.synchronize()
        want = se.istft_custom(se.apply_mask(se.stft_custom(x, c), raw, "E", True), 16000, c)
    assert float((y - want).abs().max()) == 0.0
    assert float((y2 - want).abs().max()) < 1e-5


def test_psa_loss_and_gradient(se, oref):
    """src/loss.py:32-56, the `psa` training loss (src/distrib.py:271-272)."""
    g = torch.Generator().manual_seed(4)
    shape = (3, 1, 257, 40, 2)
    enh, tgt, mix = (torch.randn(*shape, generator=g) for _ in range(3))
    er = enh.double().requires_grad_(True)
    want = oref.psa_loss_ref(er, tgt.double(), mix.double())
    (gw,) = torch.autograd.grad(want, er)
    ec = enh.cuda().requires_grad_(True)
    got = se.loss_phase_sensitive_spectral_approximation(ec, tgt.cuda(), mix.cuda())
    assert got.dim() == 0
    assert abs(float(got) - float(want)) / float(want) < 1e-4
    (gg,) = torch.autograd.grad(got, ec)
    assert rel(gg, gw) < TOL_GRAD
    assert abs(float(got) - float(oref.psa_loss_ref(enh, tgt, mix))) / float(want) < 1e-4
