"""numpy float64 restatement from first principles (TEST INFRASTRUCTURE ONLY).

No torch.stft here: explicit reflect pad, framing, window, DFT, overlap-add, envelope and the
adjoints, following SURVEY.md 8(a) rows a1-a4 / a8 (which restate src/evaluate.py:101-162,
torch/functional.py:676-680 for the pad rule, and src/model/dccrn.py:649-747).
Used for error budgeting of the fp32 paths and as a second opinion on the fp32 oracle.
"""
from __future__ import annotations

import numpy as np


def hann_periodic(win_length):
    j = np.arange(win_length, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * j / win_length)


def padded_window(n_fft, win_length):
    """torch.stft centres a short window inside n_fft: left = (n_fft - win_length)//2."""
    w = np.zeros(n_fft)
    left = (n_fft - win_length) // 2
    w[left:left + win_length] = hann_periodic(win_length)
    return w


def reflect_index(i, n):
    """index into x[0:n] for padded coordinate i (may be <0 or >=n), 'reflect' rule."""
    i = np.abs(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def frame_indices(nsample, n_fft, hop):
    """[T, n_fft] gather indices into x for centre=True reflect framing; T = 1 + N//hop."""
    nframe = 1 + nsample // hop
    pos = np.arange(nframe)[:, None] * hop + np.arange(n_fft)[None, :] - n_fft // 2
    return reflect_index(pos, nsample)


def stft(x, n_fft, hop, win_length, scale=None):
    """x [rows,N] -> X [rows,F,T] complex128, multiplied by `scale` (reference: 1/win_length)."""
    x = np.asarray(x, dtype=np.float64)
    idx = frame_indices(x.shape[-1], n_fft, hop)
    frames = x[:, idx] * padded_window(n_fft, win_length)          # [rows,T,n]
    spec = np.fft.rfft(frames, axis=-1).transpose(0, 2, 1)
    return spec * ((1.0 / win_length) if scale is None else scale)


def ola_envelope(n_fft, hop, win_length, nframe):
    w2 = padded_window(n_fft, win_length) ** 2
    env = np.zeros(n_fft + hop * (nframe - 1))
    for t in range(nframe):
        env[t * hop: t * hop + n_fft] += w2
    return env


def istft(spec, n_fft, hop, win_length, length, scale=None):
    """spec [rows,F,T] complex -> y [rows,length]; `scale` multiplies the spectrum first
    (reference: win_length).  Imag parts of DC / Nyquist are ignored (C2R)."""
    spec = np.asarray(spec, dtype=np.complex128) * (win_length if scale is None else scale)
    rows, _, nframe = spec.shape
    w = padded_window(n_fft, win_length)
    frames = np.fft.irfft(spec.transpose(0, 2, 1), n=n_fft, axis=-1) * w      # [rows,T,n]
    total = n_fft + hop * (nframe - 1)
    y = np.zeros((rows, total))
    for t in range(nframe):
        y[:, t * hop: t * hop + n_fft] += frames[:, t]
    env = ola_envelope(n_fft, hop, win_length, nframe)
    start = n_fft // 2
    end = start + length
    out = np.zeros((rows, length))
    take = min(end, total) - start
    if np.min(np.abs(env[start:start + take])) < 1e-11:
        raise RuntimeError("window overlap add min < 1e-11")
    out[:, :take] = y[:, start:start + take] / env[start:start + take]
    return out


def stft_adjoint(gspec, nsample, n_fft, hop, win_length, scale=None):
    """Gradient of sum(Re(conj(G) * X)) wrt x, X = stft(x).  gspec [rows,F,T] complex
    holds dL/dRe + i dL/dIm."""
    g = np.asarray(gspec, dtype=np.complex128) * ((1.0 / win_length) if scale is None else scale)
    rows, nf, nframe = g.shape
    h = g.copy()
    h[:, 1:nf - 1] *= 0.5
    h[:, 0] = h[:, 0].real
    h[:, nf - 1] = h[:, nf - 1].real
    frames = np.fft.irfft(h.transpose(0, 2, 1), n=n_fft, axis=-1) * n_fft * padded_window(n_fft, win_length)
    gp = np.zeros((rows, nsample + n_fft))
    for t in range(nframe):
        gp[:, t * hop: t * hop + n_fft] += frames[:, t]
    half = n_fft // 2
    gx = gp[:, half:half + nsample].copy()
    j = np.arange(half)
    np.add.at(gx, (slice(None), half - j), gp[:, j])                  # left mirror
    np.add.at(gx, (slice(None), nsample - 2 - j), gp[:, half + nsample + j])  # right mirror
    return gx


def istft_adjoint(gy, nframe, n_fft, hop, win_length, scale=None):
    """Gradient wrt spec (as dRe + i dIm) of sum(gy * istft(spec))."""
    gy = np.asarray(gy, dtype=np.float64)
    rows, length = gy.shape
    env = ola_envelope(n_fft, hop, win_length, nframe)
    total = env.shape[0]
    start = n_fft // 2
    take = min(start + length, total) - start
    gp = np.zeros((rows, total))
    gp[:, start:start + take] = gy[:, :take] / env[start:start + take]
    pos = np.arange(nframe)[:, None] * hop + np.arange(n_fft)[None, :]
    frames = gp[:, pos] * padded_window(n_fft, win_length)
    spec = np.fft.rfft(frames, axis=-1).transpose(0, 2, 1)
    c = np.full(spec.shape[1], 2.0)
    c[0] = c[-1] = 1.0
    spec = spec * c[None, :, None] / n_fft
    spec[:, 0] = spec[:, 0].real
    spec[:, -1] = spec[:, -1].real
    return spec * (win_length if scale is None else scale)


def mask_apply(spec, mask, mode, pre_tanh=False):
    """complex128 spec [...,F,T]; mask real [...,F,T] ('real') or complex ('E','C','R')."""
    if pre_tanh:
        mask = np.tanh(mask) if mode == "real" else np.tanh(mask.real) + 1j * np.tanh(mask.imag)
    if mode == "real":
        return spec * mask
    if mode == "C":
        return spec * mask
    if mode == "R":
        return spec.real * mask.real + 1j * spec.imag * mask.imag
    if mode == "E":
        mag = np.sqrt(np.abs(spec) ** 2 + 1e-8)
        ph = np.arctan2(spec.imag, spec.real)
        mm = np.abs(mask)
        mph = np.arctan2(mask.imag / (mm + 1e-8), mask.real / (mm + 1e-8))
        return np.tanh(mm) * mag * np.exp(1j * (ph + mph))
    raise ValueError(mode)


def mrstft_loss(est, ref, resolutions=((512, 128, 512), (1024, 256, 1024), (2048, 512, 2048)),
                clamp=1e-7, with_grad=False):
    """SURVEY 8(c) loss in float64; optionally the analytic gradient wrt est."""
    est = np.asarray(est, dtype=np.float64).reshape(-1, np.shape(est)[-1])
    ref = np.asarray(ref, dtype=np.float64).reshape(-1, np.shape(ref)[-1])
    total, grad = 0.0, np.zeros_like(est)
    for n_fft, hop, win in resolutions:
        sa = stft(est, n_fft, hop, win, scale=1.0)
        sb = stft(ref, n_fft, hop, win, scale=1.0)
        pa, pb = np.abs(sa) ** 2, np.abs(sb) ** 2
        a, b = np.sqrt(np.maximum(pa, clamp)), np.sqrt(np.maximum(pb, clamp))
        d2, b2 = np.sum((b - a) ** 2), np.sum(b ** 2)
        lmag = np.mean(np.abs(np.log(b) - np.log(a)))
        total += np.sqrt(d2) / np.sqrt(b2) + lmag
        if with_grad:
            dl_da = (a - b) / (np.sqrt(d2) * np.sqrt(b2)) + np.sign(np.log(a) - np.log(b)) / (a * a.size)
            gspec = np.where(pa >= clamp, dl_da / a, 0.0) * sa
            grad += stft_adjoint(gspec, est.shape[-1], n_fft, hop, win, scale=1.0)
    k = len(resolutions)
    return (total / k, grad / k) if with_grad else total / k


# ------------------------------------------------------------------ DCCRN conv transforms
def conv_stft(x, win_len, win_inc, fft_len, window):
    """zero-pad win_len-win_inc each side, frames of win_len at stride win_inc,
    x window, fft_len-point rfft of the frame zero-padded at the END.  -> [rows, 2F, T]."""
    x = np.asarray(x, dtype=np.float64)
    pad = win_len - win_inc
    xp = np.pad(x, [(0, 0), (pad, pad)])
    nframe = (xp.shape[-1] - win_len) // win_inc + 1
    pos = np.arange(nframe)[:, None] * win_inc + np.arange(win_len)[None, :]
    spec = np.fft.rfft(xp[:, pos] * window, n=fft_len, axis=-1).transpose(0, 2, 1)
    return np.concatenate([spec.real, spec.imag], axis=1)


def conv_istft(spec2, win_len, win_inc, fft_len, window, length=None):
    """SURVEY a4 closed form of the pinv basis (Sherman-Morrison parity correction)."""
    spec2 = np.asarray(spec2, dtype=np.float64)
    rows, two_f, nframe = spec2.shape
    nf = two_f // 2
    k = np.arange(nf)[:, None]
    j = np.arange(win_len)[None, :]
    ang = 2.0 * np.pi * k * j / fft_len
    re, im = spec2[:, :nf].transpose(0, 2, 1), spec2[:, nf:].transpose(0, 2, 1)   # [rows,T,F]
    v = re @ np.cos(ang) - im @ np.sin(ang)                                       # [rows,T,win]
    # pinv of K (2F x win_len, K = [cos; -sin]):  (K^T K)^-1 K^T, K^T K = (fft_len/2) I + parity blocks
    ktk = np.concatenate([np.cos(ang), -np.sin(ang)], 0)
    ktk = ktk.T @ ktk
    frames = np.linalg.solve(ktk, v.reshape(-1, win_len).T).T.reshape(rows, nframe, win_len) * window
    total = win_len + win_inc * (nframe - 1)
    y = np.zeros((rows, total))
    env = np.zeros(total)
    for t in range(nframe):
        y[:, t * win_inc: t * win_inc + win_len] += frames[:, t]
        env[t * win_inc: t * win_inc + win_len] += window ** 2
    y = y / (env + 1e-8)
    pad = win_len - win_inc
    return y[:, pad:pad + length] if length else y[:, pad:total - pad]
