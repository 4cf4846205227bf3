"""The general-geometry transform path (csrc/se_generic.cuh: any power-of-two n_fft, any hop / win_length; any DCCRN
win_len / win_inc / fft_len) under the CUDA-semantics emulator, against the float64 restatement and the torch oracle.
These are the geometries the reference's signatures accept (src/evaluate.py:101-162, src/model/dccrn.py:669-747) but its
configs do not use; the tuned engine does not compile them."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emu"))
import emu_lib as E  # noqa: E402
from oracle import spectral_np64 as o64  # noqa: E402
from oracle import spectral_oracle as oref  # noqa: E402


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)


def c2(a):
    return a[..., 0] + 1j * a[..., 1]


def r2(c):
    return np.ascontiguousarray(np.stack([c.real, c.imag], -1).astype(np.float32))


# (n_fft, hop, win_length, N): hops that do not divide n_fft, small / large n_fft (radix-2 tail when log2(n/2) is odd),
# short windows, hop = 1 sample, rows shorter than n_fft
CASES = [(256, 64, 256, 1500), (512, 160, 400, 2400), (512, 100, 512, 1777), (1024, 341, 1024, 4000), (4096, 1024, 4096, 9000),
         (64, 16, 64, 333), (16, 5, 16, 97), (8, 2, 8, 41), (128, 1, 128, 200), (512, 128, 512, 300), (8192, 2048, 8192, 9000),
         (2048, 300, 1200, 5000),
         # n_fft that is not a power of two: Bluestein (the commented CRN setting 320/160, config.yaml:78-80; 400-point front ends)
         (320, 160, 320, 3000), (400, 100, 400, 2222), (400, 160, 400, 1600), (100, 25, 100, 700), (360, 90, 300, 2000), (12, 3, 12, 50), (1000, 250, 1000, 2800), (6000, 1500, 6000, 7000), (10, 2, 10, 40)]


@pytest.mark.parametrize("n,hop,win,N", CASES)
def test_general_geometry_transforms_vs_f64(n, hop, win, N):
    rng = np.random.default_rng(n + hop + N)
    x = rng.standard_normal((2, N)).astype(np.float32)
    T, F = 1 + N // hop, n // 2 + 1
    tuned = (n in (512, 1024, 2048)) and (hop * 4 == n or hop * 2 == n)
    s = E.stft_fwd(x, n, hop, win, 1.0 / win)
    assert not np.isnan(s).any()
    assert rel(c2(s), o64.stft(x, n, hop, win)) < 2e-6
    spec = rng.standard_normal((2, F, T)) + 1j * rng.standard_normal((2, F, T))
    g = r2(spec)
    gx = E.stft_bwd(g, N, n, hop, win, 1.0 / win)
    assert not np.isnan(gx).any()
    assert rel(gx, o64.stft_adjoint(c2(g), N, n, hop, win)) < 3e-6
    base = rng.standard_normal((2, N)).astype(np.float32)
    gx2 = E.stft_bwd(g, N, n, hop, win, 1.0 / win, accumulate=True, init=base)
    assert rel(gx2 - base, gx) < 1e-5
    if tuned:
        return                                          # N < n_fft: only the adjoint takes the general path
    for length in (N, N - N // 7) + ((N + 50,) if win == n else ()):
        y = E.istft_fwd(r2(spec), length, n, hop, win, float(win))
        assert not np.isnan(y).any()
        assert rel(y, o64.istft(spec.astype(np.complex64), n, hop, win, length)) < 3e-6
        gy = rng.standard_normal((2, length)).astype(np.float32)
        gs = E.istft_bwd(gy, T, n, hop, win, float(win))
        assert not np.isnan(gs).any()
        assert rel(c2(gs), o64.istft_adjoint(gy, T, n, hop, win)) < 3e-6


@pytest.mark.parametrize("n,hop,win,N", [(256, 64, 256, 1500), (512, 160, 400, 2400), (4096, 1024, 4096, 9000)])
def test_general_geometry_matches_torch_oracle(n, hop, win, N):
    """The same path against the reference's torch.stft / torch.istft call pattern (oracle/spectral_oracle.py)."""
    import types
    cfg = types.SimpleNamespace(n_fft=n, hop_length=hop, win_length=win, center=True)
    x = torch.randn(2, 1, N, generator=torch.Generator().manual_seed(n))
    want = oref.stft_custom_ref(x, cfg)
    got = E.stft_fwd(np.ascontiguousarray(x[:, 0].numpy()), n, hop, win, 1.0 / win)
    assert rel(got, want[:, 0].numpy()) < 1e-4
    y = oref.istft_custom_ref(want, N, cfg)
    back = E.istft_fwd(np.ascontiguousarray(want[:, 0].numpy()), N, n, hop, win, float(win))
    assert rel(back, y[:, 0].numpy()) < 1e-4
    assert rel(back, x[:, 0].numpy()) < 1e-4            # analysis -> synthesis round trip


def test_general_geometry_envelope_error():
    """hop = n_fft with a Hann window: the overlap-add envelope touches zero -- torch.istft raises, so does the kernel path."""
    spec = np.zeros((1, 129, 5, 2), np.float32)
    with pytest.raises(RuntimeError, match="overlap add"):
        E.istft_fwd(spec, 900, 256, 256, 256, 256.0)


def test_general_geometry_rejects_what_it_cannot_do():
    x = np.zeros((1, 4000), np.float32)
    for n, hop, win in ((401, 100, 401), (512, 600, 512), (512, 128, 600), (16384, 4096, 16384)):
        with pytest.raises(RuntimeError):
            E.stft_fwd(x, n, hop, win, 1.0)


# (win_len, win_inc, fft_len): the reference's ConvSTFT / ConviSTFT take any of these (dccrn.py:669-747)
CONV_CASES = [(320, 160, 512), (512, 128, 512), (400, 100, 1024), (256, 64, 256), (25, 10, 32), (400, 160, 512), (401, 100, 512),
              (400, 128, 512), (400, 100, 400), (300, 75, 360)]


@pytest.mark.parametrize("wl,inc,nfft", CONV_CASES)
def test_general_geometry_dccrn_transforms(wl, inc, nfft):
    rng = np.random.default_rng(wl + inc + nfft)
    N = 20 * inc + 37
    x = rng.standard_normal((2, N)).astype(np.float32)
    window = o64.hann_periodic(wl)
    s = E.conv_stft_fwd(x, wl, inc, nfft)
    want = o64.conv_stft(x, wl, inc, nfft, window)
    assert s.shape == want.shape and not np.isnan(s).any()
    assert rel(s, want) < 2e-6
    spec = rng.standard_normal(want.shape).astype(np.float32)
    T = spec.shape[-1]
    natural = inc * (T - 1) + wl - 2 * (wl - inc)
    for out_len in (natural, natural - 11, natural + (wl - inc)):
        y = E.conv_istft_fwd(spec, out_len, wl, inc, nfft)
        assert not np.isnan(y).any()
        assert rel(y, o64.conv_istft(spec, wl, inc, nfft, window, out_len)) < 5e-6
    # against the reference's conv_transpose1d formulation and its autograd
    st = torch.from_numpy(spec).double().requires_grad_(True)
    yr = oref.conv_istft_ref(st, wl, inc, nfft, "hann", natural)
    assert rel(E.conv_istft_fwd(spec, natural, wl, inc, nfft), yr[:, 0].detach().numpy()) < 1e-5
    gy = rng.standard_normal((2, natural)).astype(np.float32)
    (gs,) = torch.autograd.grad(yr, st, torch.from_numpy(gy).double()[:, None])
    got = E.conv_istft_bwd(gy, T, wl, inc, nfft)
    assert not np.isnan(got).any()
    assert rel(got, gs.numpy()) < 1e-5


def test_general_geometry_dccrn_any_window():
    from scipy.signal import get_window
    wl, inc, nfft = 320, 80, 512
    window = get_window("hamming", wl, fftbins=True)
    wid = E.register_window(window)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 2000)).astype(np.float32)
    s = E.conv_stft_fwd_w(x, wl, inc, nfft, wid)
    assert rel(s, o64.conv_stft(x, wl, inc, nfft, window)) < 2e-6
    T = s.shape[-1]
    natural = inc * (T - 1) + wl - 2 * (wl - inc)
    y = E.conv_istft_fwd_w(s, natural, wl, inc, nfft, wid)
    assert rel(y, o64.conv_istft(s, wl, inc, nfft, window, natural)) < 5e-6


def test_general_geometry_matches_reference_run_goldens():
    """tests/golden/general_geometry.npz holds outputs of the REAL reference (stft_custom / istft_custom and their
    autograd, ConvSTFT / ConviSTFT and its autograd) at geometries outside its configs (make_golden.py::gen_general_geometry)."""
    from conftest import golden
    from scipy.signal import get_window
    g = golden("general_geometry")
    for i, (N, n, h, w) in enumerate(g["t_meta"]):
        N, n, h, w = int(N), int(n), int(h), int(w)
        k = lambda name: np.ascontiguousarray(g[f"t{i}_{name}"][:, 0])
        assert rel(E.stft_fwd(k("x"), n, h, w, 1.0 / w), k("spec")) < 1e-4
        assert rel(E.stft_bwd(k("gspec"), N, n, h, w, 1.0 / w), k("gx")) < 1e-4
        assert rel(E.istft_fwd(k("s"), N, n, h, w, float(w)), k("y")) < 1e-4
        assert rel(E.istft_bwd(k("gy"), k("s").shape[2], n, h, w, float(w)), k("gs")) < 1e-4
    for i, (N, n, h, w) in enumerate(g["n_meta"]):                      # config.center = False
        N, n, h, w = int(N), int(n), int(h), int(w)
        k = lambda name: np.ascontiguousarray(g[f"n{i}_{name}"][:, 0])
        assert rel(E.stft_nocenter_fwd(k("x"), n, h, w, 1.0 / w), k("spec")) < 1e-4
        assert rel(E.stft_nocenter_bwd(k("gspec"), N, n, h, w, 1.0 / w), k("gx")) < 1e-4
    for i, (N, wl, inc, nfft, wt) in enumerate(g["c_meta"]):
        wl, inc, nfft = int(wl), int(inc), int(nfft)
        wid = E.register_window(np.asarray(get_window("hamming" if wt else "hann", wl, fftbins=True), dtype=np.float64))
        k = lambda name: np.ascontiguousarray(g[f"c{i}_{name}"])
        assert rel(E.conv_stft_fwd_w(np.ascontiguousarray(k("x")[:, 0]), wl, inc, nfft, wid), k("spec")) < 1e-4
        y = k("y")[:, 0]
        assert rel(E.conv_istft_fwd_w(k("s"), y.shape[-1], wl, inc, nfft, wid), y) < 1e-4
        assert rel(E.conv_istft_bwd_w(np.ascontiguousarray(k("gy")[:, 0]), k("s").shape[-1], wl, inc, nfft, wid), k("gs")) < 1e-4


def test_general_geometry_row_batches(monkeypatch):
    """Synthesis scratch is bounded by processing rows in batches; force batches of one and two rows."""
    rng = np.random.default_rng(8)
    n, hop, win, N = 256, 50, 256, 1300
    T, F = 1 + N // hop, n // 2 + 1
    spec = rng.standard_normal((5, F, T)) + 1j * rng.standard_normal((5, F, T))
    want_y = o64.istft(spec.astype(np.complex64), n, hop, win, N)
    want_g = o64.stft_adjoint(spec, N, n, hop, win)
    for cap in (T * n, 2 * T * n + 5):
        monkeypatch.setenv("SE_GEN_SCRATCH_FLOATS", str(cap))
        assert rel(E.istft_fwd(r2(spec), N, n, hop, win, float(win)), want_y) < 3e-6
        assert rel(E.stft_bwd(r2(spec), N, n, hop, win, 1.0 / win), want_g) < 3e-6


@pytest.mark.parametrize("n,hop,win,N", [(512, 128, 512, 2000), (1024, 256, 1024, 5000), (256, 100, 200, 1234), (320, 160, 320, 3000)])
def test_stft_without_centre_padding(n, hop, win, N):
    """config.center = False (src/evaluate.py:116): torch.stft without padding, and its gradient, against torch."""
    x = torch.randn(2, N, generator=torch.Generator().manual_seed(n + hop), dtype=torch.float64, requires_grad=True)
    w = torch.hann_window(win, dtype=torch.float64)
    want = torch.view_as_real(torch.stft(x, n, hop, win, w, center=False, return_complex=True)) / win
    got = E.stft_nocenter_fwd(np.ascontiguousarray(x.detach().numpy().astype(np.float32)), n, hop, win, 1.0 / win)
    assert got.shape == tuple(want.shape)
    assert rel(got, want.detach().numpy()) < 2e-6
    g = torch.randn(want.shape, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    (gx,) = torch.autograd.grad(want, x, g)
    got_g = E.stft_nocenter_bwd(np.ascontiguousarray(g.numpy().astype(np.float32)), N, n, hop, win, 1.0 / win)
    assert rel(got_g, gx.numpy()) < 3e-6


def test_general_geometry_random_sweep():
    """Seeded random geometries (even n_fft 8..260, any hop <= n_fft / 2, any win_length, ragged lengths): analysis and both
    adjoints against the float64 restatement; synthesis where torch.istft would accept the envelope."""
    rng = np.random.default_rng(20261017)
    done = 0
    while done < 14:
        n = 2 * int(rng.integers(4, 131))
        hop = int(rng.integers(1, n // 2 + 1))
        win = int(rng.integers(max(2, n // 2), n + 1))
        N = int(rng.integers(n // 2 + 1, 6 * n))
        if (n in (512, 1024, 2048)) and (hop * 4 == n or hop * 2 == n):
            continue
        done += 1
        T, F = 1 + N // hop, n // 2 + 1
        x = rng.standard_normal((2, N)).astype(np.float32)
        tag = (n, hop, win, N)
        assert rel(c2(E.stft_fwd(x, n, hop, win, 1.0 / win)), o64.stft(x, n, hop, win)) < 3e-6, tag
        spec = rng.standard_normal((2, F, T)) + 1j * rng.standard_normal((2, F, T))
        assert rel(E.stft_bwd(r2(spec), N, n, hop, win, 1.0 / win), o64.stft_adjoint(spec, N, n, hop, win)) < 5e-6, tag
        try:
            want = o64.istft(spec.astype(np.complex64), n, hop, win, N)
        except RuntimeError:
            with pytest.raises(RuntimeError, match="overlap add"):
                E.istft_fwd(r2(spec), N, n, hop, win, float(win))
            continue
        assert rel(E.istft_fwd(r2(spec), N, n, hop, win, float(win)), want) < 2e-5, tag
        gy = rng.standard_normal((2, N)).astype(np.float32)
        assert rel(c2(E.istft_bwd(gy, T, n, hop, win, float(win))), o64.istft_adjoint(gy, T, n, hop, win)) < 2e-5, tag
