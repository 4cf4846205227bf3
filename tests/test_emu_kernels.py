"""Kernel sources run under the CUDA-semantics emulator (tests/cuda_emu) and compared with the
oracle -- CPU-only debugging aid for index math (framing, reflect folds, OLA carries, chunking).
The emulated library is test infrastructure; the product never loads it."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emu"))
import emu_lib as E  # noqa: E402
from oracle import spectral_np64 as o64  # noqa: E402
from oracle import spectral_oracle as oref  # noqa: E402
from conftest import golden  # noqa: E402


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)


def c2(a):
    return a[..., 0] + 1j * a[..., 1]


def r2(c):
    return np.ascontiguousarray(np.stack([c.real, c.imag], -1).astype(np.float32))


CASES = [(512, 128, 512, 2085), (1024, 256, 1024, 4351), (2048, 512, 2048, 8705), (512, 256, 512, 3000),
         (512, 128, 400, 2500), (1024, 512, 1024, 4100), (512, 128, 512, 9001)]


@pytest.mark.parametrize("n,hop,win,N", CASES)
def test_emu_transforms_vs_f64(n, hop, win, N):
    rng = np.random.default_rng(n + N)
    x = rng.standard_normal((2, N)).astype(np.float32)
    T, F = 1 + N // hop, n // 2 + 1
    s = E.stft_fwd(x, n, hop, win, 1.0 / win)
    assert rel(c2(s), o64.stft(x, n, hop, win)) < 1e-6
    spec = rng.standard_normal((2, F, T)) + 1j * rng.standard_normal((2, F, T))
    for length in (N, N - 150, N + 100):
        y = E.istft_fwd(r2(spec), length, n, hop, win, float(win))
        assert not np.isnan(y).any()
        assert rel(y, o64.istft(spec.astype(np.complex64), n, hop, win, length)) < 2e-6
        gy = rng.standard_normal((2, length)).astype(np.float32)
        gs = E.istft_bwd(gy, T, n, hop, win, float(win))
        assert rel(c2(gs), o64.istft_adjoint(gy, T, n, hop, win)) < 2e-6
    g = r2(spec)
    gx = E.stft_bwd(g, N, n, hop, win, 1.0 / win)
    assert not np.isnan(gx).any()
    assert rel(gx, o64.stft_adjoint(c2(g), N, n, hop, win)) < 2e-6
    base = rng.standard_normal((2, N)).astype(np.float32)
    gx2 = E.stft_bwd(g, N, n, hop, win, 1.0 / win, accumulate=True, init=base)
    assert rel(gx2 - base, gx) < 1e-5


@pytest.mark.parametrize("name", ["stft_n512", "stft_n1024", "stft_n2048", "stft_n512_hop256", "stft_n512_win400",
                                  "stft_n512_short_len", "stft_n512_long_len", "stft_structured"])
def test_emu_matches_reference_golden(name):
    g = golden(name)
    n, h, w, length = (int(v) for v in g["meta"])
    x = np.ascontiguousarray(g["x"].reshape(-1, g["x"].shape[-1]))
    spec = g["spec"].reshape(x.shape[0], n // 2 + 1, -1, 2)
    s = E.stft_fwd(x, n, h, w, 1.0 / w)
    assert rel(s, spec) < 1e-4                      # north_star: spectra within 1e-4 relative
    assert np.abs(s - spec).max() < 2e-6
    y = E.istft_fwd(np.ascontiguousarray(spec), length, n, h, w, float(w))
    assert rel(y, g["y"].reshape(x.shape[0], -1)) < 1e-4
    if "spec2" in g:
        y2 = E.istft_fwd(np.ascontiguousarray(g["spec2"].reshape(spec.shape)), length, n, h, w, float(w))
        assert rel(y2, g["y2"].reshape(x.shape[0], -1)) < 1e-4


@pytest.mark.parametrize("name", ["grad_n512", "grad_n1024", "grad_n512_win400"])
def test_emu_adjoints_match_reference_autograd(name):
    g = golden(name)
    n, h, w, N = (int(v) for v in g["meta"])
    rows = g["x"].shape[0] * g["x"].shape[1]
    gx = E.stft_bwd(np.ascontiguousarray(g["gspec"].reshape(rows, n // 2 + 1, -1, 2)), N, n, h, w, 1.0 / w)
    assert rel(gx, g["gx"].reshape(rows, N)) < 1e-5
    T = g["s"].shape[-2]
    gs = E.istft_bwd(np.ascontiguousarray(g["gy"].reshape(rows, N)), T, n, h, w, float(w))
    assert rel(gs, g["gs"].reshape(rows, n // 2 + 1, T, 2)) < 1e-5


@pytest.mark.parametrize("mode,name", [(0, "real"), (1, "E"), (2, "C"), (3, "R")])
@pytest.mark.parametrize("pre_tanh", [False, True])
def test_emu_masks(mode, name, pre_tanh):
    rng = np.random.default_rng(mode)
    spec = rng.standard_normal((2, 33, 7, 2)).astype(np.float32)
    spec[0, 0, 0] = 0.0                                        # atan2(0,0) corner
    m = rng.standard_normal((2, 33, 7) if mode == 0 else (2, 33, 7, 2)).astype(np.float32)
    out = E.mask_fwd(spec, m, mode, pre_tanh)
    st = torch.from_numpy(spec).double().requires_grad_(True)
    mt = torch.from_numpy(m).double().requires_grad_(True)
    o = oref.mask_apply_ref(st, mt, name, pre_tanh)
    assert rel(out, o.detach().numpy()) < 1e-6
    go = rng.standard_normal(spec.shape).astype(np.float32)
    go[0, 0, 0] = 0.0
    gs_, gm_ = torch.autograd.grad(o, (st, mt), torch.from_numpy(go).double())
    gm, gs = E.mask_bwd(spec, m, go, mode, pre_tanh)
    assert rel(gm, gm_.numpy()) < 1e-5
    assert rel(gs, np.nan_to_num(gs_.numpy())) < 1e-5


@pytest.mark.parametrize("pre_tanh", [False, True])
def test_emu_polar_mask_elementwise_accuracy_over_magnitude_ranges(pre_tanh):
    """'E' mask per element (not max-norm): |m| from 1e-6 to 60 (series / rational / saturated tanh), |x| from 0
    and 1e-20 (unit-vector path, like zero-filled segments) to 1e3."""
    rng = np.random.default_rng(11)
    mags_m = np.array([1e-6, 1e-3, 0.05, 0.0999, 0.1001, 0.3, 1.0, 3.0, 7.9, 8.0, 60.0])
    mags_x = np.array([0.0, 1e-20, 1e-14, 1e-6, 1e-4, 1e-2, 1.0, 1e3])
    ph = rng.uniform(0, 2 * np.pi, (2, mags_x.size, mags_m.size))
    x = (mags_x[:, None] * np.exp(1j * ph[0]))
    m = (mags_m[None, :] * np.exp(1j * ph[1]))
    spec = np.ascontiguousarray(np.stack([x.real, x.imag], -1)[None].astype(np.float32))
    mask = np.ascontiguousarray(np.stack([m.real, m.imag], -1)[None].astype(np.float32))
    out = E.mask_fwd(spec, mask, 1, pre_tanh)
    st = torch.from_numpy(spec).double()
    mt = torch.from_numpy(mask).double().requires_grad_(True)
    o = oref.mask_apply_ref(st, mt, "E", pre_tanh)
    want = o.detach().numpy()
    mag = np.sqrt((want ** 2).sum(-1, keepdims=True))
    assert not np.isnan(out).any()
    assert (np.abs(out - want) <= 3e-6 * mag + 1e-30).all()
    go = rng.standard_normal(spec.shape).astype(np.float32)
    (gm_,) = torch.autograd.grad(o, mt, torch.from_numpy(go).double())
    gm, _ = E.mask_bwd(spec, mask, go, 1, pre_tanh)
    gm_ = gm_.numpy()
    gmag = np.sqrt((gm_ ** 2).sum(-1, keepdims=True))
    live = np.broadcast_to(mags_x[None, :, None, None] > 0, gm.shape)      # atan2'(0,0) is undefined in the reference
    assert not np.isnan(gm).any()
    # past tanh's saturation the gradient is a difference of nearly equal terms: bound it against the scale of
    # the terms (|x| |go|) as well
    scale = np.sqrt((spec.astype(np.float64) ** 2).sum(-1, keepdims=True) + 1e-8) * np.sqrt((go.astype(np.float64) ** 2).sum(-1, keepdims=True))
    assert (np.abs(gm - gm_)[live] <= (1e-5 * gmag + 1e-6 * scale + 1e-30)[np.broadcast_to(live[..., :1], gmag.shape)].repeat(2)).all()


def test_emu_mask_matches_dcunet_and_dccrn_golden():
    g = golden("dcunet_mask_E")
    out = E.mask_fwd(np.ascontiguousarray(g["spec"]), np.ascontiguousarray(g["raw_mask"]), 1, True)
    assert rel(out, g["out"]) < 1e-5
    for mode, code in (("E", 1), ("C", 2), ("R", 3)):
        g = golden(f"dccrn_mask_{mode}")
        nf = g["specs"].shape[1] // 2
        spec = np.ascontiguousarray(np.stack([g["specs"][:, :nf], g["specs"][:, nf:]], -1))
        mask = np.ascontiguousarray(np.stack([g["mask_re"], g["mask_im"]], -1))
        out = E.mask_fwd(spec, mask, code, False)
        assert rel(np.concatenate([out[..., 0], out[..., 1]], 1), g["out_spec"]) < 1e-5


@pytest.mark.parametrize("mode,code", [("E", 1), ("C", 2), ("R", 3)])
def test_emu_planar_dccrn_mask(mode, code):
    """DCCRN-layout mask kernels (se_mask_planar_fwd/bwd): the reference's golden vectors, and forward + all three
    gradients against the interleaved kernels on a shape whose plane is odd (row bases lose 8-byte alignment)."""
    g = golden(f"dccrn_mask_{mode}")
    specs, mre, mim = (np.ascontiguousarray(g[k]) for k in ("specs", "mask_re", "mask_im"))
    assert rel(E.mask_planar_fwd(specs, mre, mim, code), g["out_spec"]) < 1e-5
    rng = np.random.default_rng(code)
    nf, nt = 9, 7
    specs = rng.standard_normal((3, 2 * nf, nt)).astype(np.float32)
    specs[1, 0, 0] = specs[1, nf, 0] = 0.0                     # atan2(0,0) corner
    mre, mim = (rng.standard_normal((3, nf, nt)).astype(np.float32) for _ in range(2))
    go = rng.standard_normal(specs.shape).astype(np.float32)
    inter = lambda a: np.ascontiguousarray(np.stack([a[:, :nf], a[:, nf:]], -1))
    planar = lambda a: np.concatenate([a[..., 0], a[..., 1]], 1)
    spec_i, mask_i = inter(specs), np.ascontiguousarray(np.stack([mre, mim], -1))
    assert np.array_equal(E.mask_planar_fwd(specs, mre, mim, code), planar(E.mask_fwd(spec_i, mask_i, code, False)))
    gm_i, gs_i = E.mask_bwd(spec_i, mask_i, inter(go), code, False)
    gre, gim, gs = E.mask_planar_bwd(specs, mre, mim, go, code)
    assert np.array_equal(gre, gm_i[..., 0]) and np.array_equal(gim, gm_i[..., 1]) and np.array_equal(gs, planar(gs_i))
    gre2, gim2, none = E.mask_planar_bwd(specs, mre, mim, go, code, want_gspec=False)
    assert none is None and np.array_equal(gre2, gre) and np.array_equal(gim2, gim)


@pytest.mark.parametrize("groups", ["2", "3"])
def test_emu_fused_dccrn_tail_multi_group_and_natural_length(groups, monkeypatch):
    """The fused DCCRN tail with forced multi-group chunks (ring carry across groups) and with `length=None`
    (natural length, both pads dropped), against the two-stage kernels planned the default way."""
    rng = np.random.default_rng(7)
    T = 61
    natural = 100 * (T - 1) + 400 - 2 * 300
    spec = rng.standard_normal((1, 514, T)).astype(np.float32)
    mre, mim = (rng.standard_normal((1, 257, T)).astype(np.float32) for _ in range(2))
    want = E.conv_istft_fwd(E.mask_planar_fwd(spec, mre, mim, 1), natural, 400, 100, 512)
    gy = rng.standard_normal(want.shape).astype(np.float32)
    wre, wim, _ = E.mask_planar_bwd(spec, mre, mim, E.conv_istft_bwd(gy, T, 400, 100, 512), 1)
    monkeypatch.setenv("SE_FORCE_GROUPS", groups)
    y = E.conv_mask_istft_fwd(spec, mre, mim, natural, 400, 100, 512, 1)
    gre, gim = E.conv_mask_istft_bwd(gy, spec, mre, mim, 400, 100, 512, 1)
    assert rel(y, want) < 1e-6 and rel(gre, wre) < 1e-6 and rel(gim, wim) < 1e-6


@pytest.mark.parametrize("mode,code", [("E", 1), ("C", 2), ("R", 3)])
def test_emu_fused_dccrn_tail(mode, code):
    """se_conv_mask_istft_fwd/bwd = ConviSTFT(mask tail) and its gradient wrt the two mask planes, against the two-stage
    emulated kernels (planar mask, then ConviSTFT / its adjoint, then the mask adjoint)."""
    rng = np.random.default_rng(40 + code)
    T, out_len = 37, 3000
    spec = rng.standard_normal((2, 514, T)).astype(np.float32)
    mre, mim = (rng.standard_normal((2, 257, T)).astype(np.float32) for _ in range(2))
    mre[:, 0] = mim[:, 0] = 0.0                                  # the model pads the masks at DC (dccrn.py:200-201)
    y = E.conv_mask_istft_fwd(spec, mre, mim, out_len, 400, 100, 512, code)
    masked = E.mask_planar_fwd(spec, mre, mim, code)
    want = E.conv_istft_fwd(masked, out_len, 400, 100, 512)
    assert not np.isnan(y).any() and rel(y, want) < 1e-6
    gy = rng.standard_normal(y.shape).astype(np.float32)
    gre, gim = E.conv_mask_istft_bwd(gy, spec, mre, mim, 400, 100, 512, code)
    gmasked = E.conv_istft_bwd(gy, T, 400, 100, 512)
    wre, wim, _ = E.mask_planar_bwd(spec, mre, mim, gmasked, code)
    assert not (np.isnan(gre).any() or np.isnan(gim).any())
    assert rel(gre, wre) < 1e-6 and rel(gim, wim) < 1e-6


def test_emu_mrstft_loss_and_grad():
    rng = np.random.default_rng(1)
    N = 5000
    ref = rng.standard_normal((2, N)).astype(np.float32)
    est = (ref + 0.1 * rng.standard_normal((2, N))).astype(np.float32)
    sums, loss = E.mrstft_fwd(est, ref)
    l64, g64 = o64.mrstft_loss(est, ref, with_grad=True)
    assert abs(loss - l64) / l64 < 1e-5
    parts = np.array(oref.mrstft_partials_ref(torch.from_numpy(est)[:, None], torch.from_numpy(ref)[:, None]))
    assert rel(sums.reshape(3, 3), parts[:, :3]) < 1e-5
    g = E.mrstft_bwd(est, ref, sums, 0.5)
    assert not np.isnan(g).any()
    assert rel(g, 0.5 * g64) < 1e-3                 # north_star: loss and gradients within 1e-3


@pytest.mark.parametrize("name", ["conv_a", "conv_b", "conv_c"])
def test_emu_conv_stft_matches_reference_golden(name):
    g = golden(name)
    wl, inc, nfft, _ = (int(v) for v in g["meta"])
    s = E.conv_stft_fwd(np.ascontiguousarray(g["x"][:, 0]), wl, inc, nfft)
    assert s.shape == g["spec"].shape
    assert rel(s, g["spec"]) < 1e-5


@pytest.mark.parametrize("name", ["conv_a", "conv_b", "conv_c"])
def test_emu_conv_istft_matches_reference_golden_and_autograd(name):
    g = golden(name)
    wl, inc, nfft, length = (int(v) for v in g["meta"])
    for ks, ky in (("spec", "y"), ("spec2", "y2")):
        y = E.conv_istft_fwd(np.ascontiguousarray(g[ks]), g[ky].shape[-1], wl, inc, nfft)
        assert not np.isnan(y).any()
        assert rel(y, g[ky][:, 0]) < 1e-5
    st = torch.from_numpy(g["spec2"]).double().requires_grad_(True)
    yr = oref.conv_istft_ref(st, wl, inc, nfft, "hann", None if length < 0 else length)
    gy = np.random.default_rng(0).standard_normal(yr.shape).astype(np.float32)
    (gs,) = torch.autograd.grad(yr, st, torch.from_numpy(gy).double())
    got = E.conv_istft_bwd(np.ascontiguousarray(gy[:, 0]), st.shape[-1], wl, inc, nfft)
    assert rel(got, gs.numpy()) < 1e-5


@pytest.mark.parametrize("n,hop,win,N", [(512, 128, 512, 2085), (1024, 256, 1024, 4351), (512, 256, 512, 3000),
                                         (512, 128, 400, 2500), (2048, 512, 2048, 8705)])
@pytest.mark.parametrize("mode,name", [(0, "real"), (1, "E"), (2, "C")])
def test_emu_fused_enhance(n, hop, win, N, mode, name):
    import types
    rng = np.random.default_rng(n + mode)
    pre_tanh = mode == 1
    F, T = n // 2 + 1, 1 + N // hop
    cfg = types.SimpleNamespace(n_fft=n, hop_length=hop, win_length=win, center=True)
    x = rng.standard_normal((2, N)).astype(np.float32)
    m = rng.standard_normal((2, F, T) if mode == 0 else (2, F, T, 2)).astype(np.float32)
    y = E.enhance_fwd(x, m, n, hop, win, mode, pre_tanh)
    xt = torch.from_numpy(x)[:, None].double()
    mt = torch.from_numpy(m)[:, None].double().requires_grad_(True)
    yr = oref.istft_custom_ref(oref.mask_apply_ref(oref.stft_custom_ref(xt, cfg), mt, name, pre_tanh), N, cfg)
    assert not np.isnan(y).any()
    assert rel(y, yr.detach().numpy()[:, 0]) < 2e-6
    if n <= 1024:
        gy = rng.standard_normal((2, N)).astype(np.float32)
        (gm_ref,) = torch.autograd.grad(yr, mt, torch.from_numpy(gy)[:, None].double())
        gm = E.enhance_bwd(gy, x, m, n, hop, win, mode, pre_tanh)
        assert rel(gm, gm_ref.numpy()[:, 0]) < 2e-6


@pytest.mark.parametrize("n,hop,win,N,length", [(512, 128, 512, 2085, 2085), (1024, 256, 1024, 4351, 4000),
                                                (512, 256, 400, 3000, 3000), (2048, 512, 2048, 8705, 8705),
                                                (1024, 512, 1024, 5000, 5200)])
@pytest.mark.parametrize("mode,name", [(0, "real"), (1, "E"), (2, "C"), (3, "R")])
def test_emu_mask_istft(n, hop, win, N, length, mode, name):
    """Model tail + istft_custom in one launch, and its backward to the raw mask (spectrum never written)."""
    rng = np.random.default_rng(7 * n + mode)
    pre_tanh = mode in (1, 3)
    F, T = n // 2 + 1, 1 + N // hop
    if length > hop * (T - 1) + n // 2:
        length = hop * (T - 1)
    cfg = oref.make_config(n, hop, win)
    spec = rng.standard_normal((2, F, T, 2)).astype(np.float32)
    m = rng.standard_normal((2, F, T) if mode == 0 else (2, F, T, 2)).astype(np.float32)
    y = E.mask_istft_fwd(spec, m, length, n, hop, win, float(win), mode, pre_tanh)
    st = torch.from_numpy(spec)[:, None].double()
    mt = torch.from_numpy(m)[:, None].double().requires_grad_(True)
    yr = oref.istft_custom_ref(oref.mask_apply_ref(st, mt, name, pre_tanh), length, cfg)
    assert not np.isnan(y).any()
    assert rel(y, yr.detach().numpy()[:, 0]) < 2e-6
    gy = rng.standard_normal((2, length)).astype(np.float32)
    (gm_ref,) = torch.autograd.grad(yr, mt, torch.from_numpy(gy)[:, None].double())
    gm = E.mask_istft_bwd(gy, spec, m, n, hop, win, float(win), mode, pre_tanh)
    assert not np.isnan(gm).any()
    assert rel(gm, gm_ref.numpy()[:, 0]) < 2e-6


def test_emu_segment_stft_matches_reference_segmenting():
    """evaluate()'s zero-filled overlapping segments + STFT in one launch (src/evaluate.py:29-39,164-183)."""
    g = golden("segments")
    nfeat, stride = (int(v) for v in g["meta"])
    wav = np.ascontiguousarray(g["wav"][0])                       # [C, L]
    nseg = g["seg"].shape[0]
    got = E.stft_segments_fwd(wav, nseg, stride, nfeat, 512, 128, 512, 1.0 / 512)
    cfg = oref.make_config(512, 128, 512)
    want = oref.stft_custom_ref(torch.from_numpy(g["seg"]).reshape(nseg, wav.shape[0], nfeat), cfg).numpy()
    assert rel(got.reshape(want.shape), want) < 1e-6


@pytest.mark.parametrize("groups", ["2", "3"])
def test_emu_multi_group_chunks_carry_path(groups, monkeypatch):
    """Force g groups per chunk so the OLA carry across groups (registers) is exercised on small inputs."""
    monkeypatch.setenv("SE_FORCE_GROUPS", groups)
    rng = np.random.default_rng(int(groups))
    n, hop, win, N = 512, 128, 512, 9001
    T, F = 1 + N // hop, n // 2 + 1
    spec = rng.standard_normal((2, F, T)) + 1j * rng.standard_normal((2, F, T))
    y = E.istft_fwd(r2(spec), N, n, hop, win, float(win))
    assert rel(y, o64.istft(spec.astype(np.complex64), n, hop, win, N)) < 2e-6
    gx = E.stft_bwd(r2(spec), N, n, hop, win, 1.0 / win)
    assert rel(gx, o64.stft_adjoint(c2(r2(spec)), N, n, hop, win)) < 2e-6
    ref = rng.standard_normal((2, N)).astype(np.float32)
    est = (ref + 0.1 * rng.standard_normal((2, N))).astype(np.float32)
    sums, loss = E.mrstft_fwd(est, ref)
    l64, g64 = o64.mrstft_loss(est, ref, with_grad=True)
    g = E.mrstft_bwd(est, ref, sums, 1.0)
    assert not np.isnan(g).any()
    assert rel(g, g64) < 2e-3
    x = rng.standard_normal((2, N)).astype(np.float32)
    m = rng.standard_normal((2, F, T, 2)).astype(np.float32)
    import types
    cfg = types.SimpleNamespace(n_fft=n, hop_length=hop, win_length=win, center=True)
    yy = E.enhance_fwd(x, m, n, hop, win, 2, False)
    yr = oref.istft_custom_ref(oref.mask_apply_ref(oref.stft_custom_ref(torch.from_numpy(x)[:, None].double(), cfg),
                                                   torch.from_numpy(m)[:, None].double(), "C"), N, cfg)
    assert rel(yy, yr.numpy()[:, 0]) < 2e-6


@pytest.mark.parametrize("kind,fn", [(0, torch.nn.functional.mse_loss), (1, torch.nn.functional.l1_loss)])
def test_emu_fused_spectral_loss(kind, fn):
    """loss_function(enhanced, stft_custom(sources)) with the target spectrum kept in registers (8f-2)."""
    rng = np.random.default_rng(kind)
    n, hop, win, N = 512, 128, 512, 3001
    cfg = oref.make_config(n, hop, win)
    tgt = rng.standard_normal((2, N)).astype(np.float32)
    tspec = oref.stft_custom_ref(torch.from_numpy(tgt)[:, None], cfg)[:, 0]
    enh = (tspec.numpy() + 0.01 * rng.standard_normal(tspec.shape)).astype(np.float32)
    loss, g = E.spectral_loss(np.ascontiguousarray(enh), tgt, n, hop, win, kind, 0.5)
    et = torch.from_numpy(enh).double().requires_grad_(True)
    want = fn(et, tspec.double())
    (gw,) = torch.autograd.grad(0.5 * want, et)
    assert abs(loss - float(want)) / float(want) < 1e-5
    assert not np.isnan(g).any()
    if kind == 0:
        assert rel(g, gw.numpy()) < 1e-4
    else:                                   # sign() flips only where |e - s| is at round-off level
        same = np.sign(g) == np.sign(gw.numpy())
        assert same.mean() > 0.999


@pytest.mark.parametrize("n,hop", [(512, 128), (512, 256), (1024, 256)])
def test_emu_edge_lengths(n, hop):
    """Lengths around every seam: the reflect limit (N = n/2+1), N = n, hop multiples +-1, and the
    13/16-block chunk boundaries; iSTFT with shorter / longer `length`; all four transforms."""
    rng = np.random.default_rng(n + hop)
    F = n // 2 + 1
    for N in (n // 2 + 1, n - 1, n, n + hop - 1, 13 * hop, 13 * hop + 1, 16 * hop - 1, 16 * hop + 1, 29 * hop + 5):
        x = rng.standard_normal((1, N)).astype(np.float32)
        T = 1 + N // hop
        assert rel(c2(E.stft_fwd(x, n, hop, n, 1.0 / n)), o64.stft(x, n, hop, n)) < 2e-6, N
        spec = rng.standard_normal((1, F, T)) + 1j * rng.standard_normal((1, F, T))
        for length in (N, max(1, N - 77), N + 50):
            y = E.istft_fwd(r2(spec), length, n, hop, n, float(n))
            assert not np.isnan(y).any(), (N, length)
            assert rel(y, o64.istft(spec.astype(np.complex64), n, hop, n, length)) < 5e-6, (N, length)
            gy = rng.standard_normal((1, length)).astype(np.float32)
            gs = E.istft_bwd(gy, T, n, hop, n, float(n))
            assert rel(c2(gs), o64.istft_adjoint(gy, T, n, hop, n)) < 5e-6, (N, length)
        if N >= n:
            gx = E.stft_bwd(r2(spec), N, n, hop, n, 1.0 / n)
            assert not np.isnan(gx).any(), N
            assert rel(gx, o64.stft_adjoint(c2(r2(spec)), N, n, hop, n)) < 5e-6, N


@pytest.mark.parametrize("tag", ["half", "quarter", "coprime", "gap", "abut"])
def test_emu_overlap_and_add_matches_reference_golden_and_autograd(tag):
    """Conv-TasNet decoder tail (src/model/conv_tasnet.py:11-31): bit-exact (same summation order); the backward
    is the exact gather."""
    g = golden("tasnet_metric")
    sig, step = g[f"sig_{tag}"], int(g[f"step_{tag}"])
    frames, length = sig.shape[-2:]
    rows = np.ascontiguousarray(sig.reshape(-1, frames, length))
    out = E.overlap_add_fwd(rows, step)
    assert np.array_equal(out.reshape(g[f"out_{tag}"].shape), g[f"out_{tag}"])
    rng = np.random.default_rng(3)
    go = rng.standard_normal(out.shape).astype(np.float32)
    st = torch.from_numpy(rows).requires_grad_(True)
    (gw,) = torch.autograd.grad(oref.overlap_and_add_ref(st, step), st, torch.from_numpy(go))
    assert np.array_equal(E.overlap_add_bwd(go, frames, length, step), gw.numpy())


# ---------------------------------------------------------------- signal-pair engine (se_fft2.cuh / se_kernels2.cuh)
@pytest.mark.parametrize("rows,N,groups", [(3, 6001, None), (1, 4100, "2"), (2, 9000, "3")])
def test_emu_pair_engine_loss_odd_rows_ragged_and_carry(rows, N, groups, monkeypatch):
    """Odd row counts (the last row is paired with itself), lengths that are no multiple of any hop, and
    multi-group chunks (OLA carry across 8-frame groups) through the pair-engine loss kernels; the scalar
    engine (SE_ENGINE=1) must agree with it to fp32 round-off."""
    if groups:
        monkeypatch.setenv("SE_FORCE_GROUPS", groups)
    rng = np.random.default_rng(rows * 7 + N)
    ref = rng.standard_normal((rows, N)).astype(np.float32)
    est = (ref + 0.1 * rng.standard_normal((rows, N))).astype(np.float32)
    l64, g64 = o64.mrstft_loss(est, ref, with_grad=True)
    out = {}
    for eng in ("2", "1"):
        monkeypatch.setenv("SE_ENGINE", eng)
        sums, loss = E.mrstft_fwd(est, ref)
        g = E.mrstft_bwd(est, ref, sums, 1.0)
        assert not np.isnan(g).any()
        assert abs(loss - l64) / l64 < 1e-5
        assert rel(g, g64) < 1e-3
        out[eng] = (sums.copy(), g.copy())
    assert rel(out["2"][0], out["1"][0]) < 1e-5
    assert rel(out["2"][1], out["1"][1]) < 1e-3      # both sit within 1e-3 of float64; the gradient is ill-conditioned


@pytest.mark.parametrize("tag", ["zscore", "plain", "n1024"])
def test_emu_evaluate_fused_stats_segments_stitch_match_reference_golden(tag):
    """evaluate(model=None) as three launches -- row statistics, segment STFT with the z-score folded into the fill,
    iSTFT that synthesises only the kept frames and writes the stitched, de-normalised clip -- against the output of the
    REAL reference's evaluate() (tests/golden/evaluate.npz) and, with a toy spectral mask between the two transforms,
    against the oracle restatement of the same flow."""
    gd = golden("evaluate")
    n_fft, hop, nfeat, zscore = (int(v) for v in gd[f"meta_{tag}"])
    mix = gd[f"mix_{tag}"]
    nb, nc, length = mix.shape
    wav = np.ascontiguousarray(mix.reshape(nb * nc, length))
    stride = n_fft
    rem = (length - nfeat) % stride
    nseg = (length + (stride - rem if rem else 0) - nfeat) // stride + 1
    stats = E.row_stats(wav) if zscore else None
    if zscore:
        t = torch.from_numpy(wav)
        assert rel(stats[:, 0], t.mean(-1).numpy()) < 1e-5 and rel(stats[:, 2], (t.std(-1) + 1e-9).numpy()) < 1e-6
    spec = E.stft_segments_norm_fwd(wav, stats, nseg, stride, nfeat, n_fft, hop, n_fft, 1.0 / n_fft)
    assert not np.isnan(spec).any()
    out = E.istft_stitch_fwd(spec, stats, nseg, nb * nc, nfeat, stride, length, n_fft, hop, n_fft, float(n_fft))
    assert not np.isnan(out).any()
    assert rel(out.reshape(mix.shape), gd[f"enh_{tag}"]) < 1e-5
    # a "model" between the transforms: fixed smooth real mask over frequency
    import types
    f = np.linspace(0.2, 1.0, n_fft // 2 + 1, dtype=np.float32)[:, None, None]
    out2 = E.istft_stitch_fwd(np.ascontiguousarray(spec * f), stats, nseg, nb * nc, nfeat, stride, length, n_fft, hop, n_fft, float(n_fft))
    conf = types.SimpleNamespace(dset=types.SimpleNamespace(norm="z-score" if zscore else "none", sample_rate=16000),
                                 model=types.SimpleNamespace(name="dnn", n_fft=n_fft, hop_length=hop, win_length=n_fft, center=True,
                                                             segment=nfeat / 16000.0))
    want = oref.evaluate_ref(torch.from_numpy(mix), lambda s: s * torch.from_numpy(f), conf).numpy()
    assert rel(out2.reshape(mix.shape), want) < 1e-5


@pytest.mark.parametrize("win_type", ["hamming", "blackman", None])
def test_emu_dccrn_transforms_with_any_scipy_window(win_type):
    """ConvSTFT / ConviSTFT with the window types the reference accepts (src/model/dccrn.py:651-655; 'hamming' is the
    constructors' default): values registered once, same kernels; against the oracle's dense conv / conv_transpose."""
    from scipy.signal import get_window
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 3000)).astype(np.float32)
    w = np.ones(400) if win_type is None else get_window(win_type, 400, fftbins=True)
    wid = E.register_window(w)
    assert E.register_window(w) == wid                                   # identical values share an id
    spec = E.conv_stft_fwd_w(x, 400, 100, 512, wid)
    want = oref.conv_stft_ref(torch.from_numpy(x), 400, 100, 512, win_type).numpy()
    assert rel(spec, want) < 1e-5
    y = E.conv_istft_fwd_w(spec, 3000, 400, 100, 512, wid)
    ywant = oref.conv_istft_ref(torch.from_numpy(spec), 400, 100, 512, win_type, length=3000).numpy()[:, 0]
    assert rel(y, ywant) < 1e-5
    gy = rng.standard_normal(y.shape).astype(np.float32)
    st = torch.from_numpy(spec).requires_grad_(True)
    (gwant,) = torch.autograd.grad(oref.conv_istft_ref(st, 400, 100, 512, win_type, length=3000), st, torch.from_numpy(gy)[:, None])
    assert rel(E.conv_istft_bwd_w(gy, spec.shape[-1], 400, 100, 512, wid), gwant.numpy()) < 1e-5


def test_emu_dccrn_polar_feature_ops():
    """ConvSTFT(feature_type='real') epilogue and ConviSTFT(inputs, phase) prologue as single launches, and the
    prologue's gradient, against torch."""
    rng = np.random.default_rng(6)
    spec = rng.standard_normal((2, 514, 9)).astype(np.float32)
    mags, phase, back = E.polar_round_trip(spec)
    t = torch.from_numpy(spec)
    assert rel(mags, torch.sqrt(t[:, :257] ** 2 + t[:, 257:] ** 2).numpy()) < 1e-6
    assert rel(phase, torch.atan2(t[:, 257:], t[:, :257]).numpy()) < 1e-6
    assert rel(back, spec) < 1e-6
    g = rng.standard_normal(spec.shape).astype(np.float32)
    m, p = torch.from_numpy(mags).requires_grad_(True), torch.from_numpy(phase).requires_grad_(True)
    gm, gp = torch.autograd.grad(torch.cat([m * torch.cos(p), m * torch.sin(p)], 1), (m, p), torch.from_numpy(g))
    got_m, got_p = E.planar_from_polar_bwd(mags, phase, g)
    assert rel(got_m, gm.numpy()) < 1e-6 and rel(got_p, gp.numpy()) < 1e-6


@pytest.mark.parametrize("n,hop,N", [(512, 128, 3001), (1024, 256, 6100)])
def test_emu_two_pass_engine_matches_the_default_engine(n, hop, N, monkeypatch):
    """SE_ENGINE=3 (csrc/se_fft3.cuh: radix-R1 x radix-16, one intermediate shared-memory round trip) against the
    three-pass engine and the float64 oracle: STFT, iSTFT with a short length, both adjoints."""
    rng = np.random.default_rng(n + N)
    x = rng.standard_normal((3, N)).astype(np.float32)
    out = {}
    for eng in ("1", "3"):
        monkeypatch.setenv("SE_ENGINE", eng)
        spec = E.stft_fwd(x, n, hop, n, 1.0 / n)
        y = E.istft_fwd(spec, N - 77, n, hop, n, float(n))
        gx = E.stft_bwd(spec, N, n, hop, n, 1.0 / n)
        gs = E.istft_bwd(y, spec.shape[2], n, hop, n, float(n))
        assert not any(np.isnan(v).any() for v in (spec, y, gx, gs))
        out[eng] = (spec, y, gx, gs)
    want = o64.stft(x.astype(np.float64), n, hop, n)                  # scaled by 1 / win_length like stft_custom
    assert rel(c2(out["3"][0]), want) < 1e-6
    for a, b in zip(out["3"], out["1"]):
        assert rel(a, b) < 2e-6


@pytest.mark.parametrize("n,hop,nfeat,length,zscore", [(512, 128, 8192, 20000, True), (512, 128, 6144, 15001, False),
                                                       (1024, 256, 8192, 17000, True)])
def test_emu_shared_frame_segment_stft_equals_per_segment_transform(n, hop, nfeat, length, zscore):
    """evaluate()'s segment STFT with the frames shared between overlapping segments transformed once at clip level
    (clip transform + boundary groups + gather) is BIT-identical to transforming every frame of every segment -- including
    the zero-filled tail of the last segments and the z-score folded into both fills."""
    rng = np.random.default_rng(n + length)
    wav = (0.3 * rng.standard_normal((2, length)) + 0.05).astype(np.float32)
    stride = n
    rem = (length - nfeat) % stride
    nseg = (length + (stride - rem if rem else 0) - nfeat) // stride + 1
    stats = E.row_stats(wav) if zscore else None
    a = E.stft_segments_norm_fwd(wav, stats, nseg, stride, nfeat, n, hop, n, 1.0 / n)
    b = E.stft_segments_shared_fwd(wav, stats, nseg, stride, nfeat, n, hop, n, 1.0 / n)
    assert not np.isnan(b).any()
    assert np.array_equal(a, b)


def test_emu_mrstft_forward_with_fused_value_matches_two_launch_form():
    """se_mrstft_loss_fwd_value (one reduction CTA that also writes the loss) against se_mrstft_loss_fwd +
    se_mrstft_loss_value: same sums bit for bit, same loss."""
    import ctypes
    rng = np.random.default_rng(3)
    ref = rng.standard_normal((3, 5000)).astype(np.float32)
    est = (ref + 0.1 * rng.standard_normal((3, 5000))).astype(np.float32)
    sums, loss = E.mrstft_fwd(est, ref)
    L = E.lib()
    ws = np.zeros(L.se_mrstft_workspace_bytes(3, 5000) // 8 + 1, np.float64)
    sums2, loss2 = np.zeros(9, np.float64), np.zeros(1, np.float32)
    E.check(L.se_mrstft_loss_fwd_value(E.ptr(est), E.ptr(ref), ctypes.c_int64(3), ctypes.c_int64(5000), E.ptr(sums2), E.ptr(loss2),
                                       E.ptr(ws), None))
    assert np.array_equal(sums, sums2)
    assert abs(float(loss2[0]) - float(loss)) <= 1e-7 * abs(float(loss))
