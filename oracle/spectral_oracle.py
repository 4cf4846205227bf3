"""torch-CPU fp32 restatement of the reference's spectral hot path (TEST INFRASTRUCTURE ONLY).

Every function names the reference lines it restates (paths relative to /root/reference).
The arithmetic itself lives in PyTorch (third-party, reference pin torch==1.7.1+cu110,
README.md:106; here 2.11.0+cu128), so the restatement calls the same library entry points
with the reference's argument choices.  See oracle/__init__.py for who may import this.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

MRSTFT_RESOLUTIONS = ((512, 128, 512), (1024, 256, 1024), (2048, 512, 2048))
MRSTFT_CLAMP = 1e-7


def make_config(n_fft, hop_length=None, win_length=None, center=True):
    """Duck-typed stand-in for the reference's config object (src/utils.py:149-165)."""
    hop_length = n_fft // 4 if hop_length is None else hop_length
    win_length = n_fft if win_length is None else win_length
    return SimpleNamespace(n_fft=n_fft, hop_length=hop_length, win_length=win_length, center=center)


def _lead_and_last(t, n_tail):
    lead = tuple(t.shape[: t.dim() - n_tail])
    return lead, tuple(t.shape[t.dim() - n_tail:])


# ----------------------------------------------------------------------------- a1 / a2
def stft_custom_ref(wave: torch.Tensor, config) -> torch.Tensor:
    """src/evaluate.py:101-128.  [B,(S,)C,N] -> [B,(S,)C,F,T,2], spectrum / win_length.

    Flatten leading dims (:107-108), torch.stft with periodic Hann of win_length, reflect
    centring, one-sided, unnormalised (:109-119), divide by win_length (:120), restore dims.
    """
    if wave.dim() not in (3, 4):
        raise ValueError("stft_custom takes [B,C,N] or [B,S,C,N]")
    lead, (nsample,) = _lead_and_last(wave, 1)
    rows = wave.contiguous().reshape(-1, nsample)
    win = torch.hann_window(config.win_length, dtype=wave.dtype, device=wave.device)
    spec = torch.stft(rows, config.n_fft, hop_length=config.hop_length,
                      win_length=config.win_length, window=win, center=config.center,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    spec = torch.view_as_real(spec) / config.win_length
    return spec.reshape(*lead, *spec.shape[1:])


def istft_custom_ref(spec: torch.Tensor, length, config) -> torch.Tensor:
    """src/evaluate.py:130-162.  [B,(S,)C,F,T,2] -> [B,(S,)C,length].

    Undo the 1/win_length scale (:131), repack to complex (:138-140), torch.istft with the same
    window/centre choices and ``length`` (:142-153).
    """
    if spec.dim() not in (5, 6):
        raise ValueError("istft_custom takes [B,C,F,T,2] or [B,S,C,F,T,2]")
    lead, (nf, nt, two) = _lead_and_last(spec, 3)
    assert two == 2
    flat = (spec * config.win_length).contiguous().reshape(-1, nf, nt, 2)
    cplx = torch.complex(flat[..., 0], flat[..., 1])
    win = torch.hann_window(config.win_length, dtype=spec.dtype, device=spec.device)
    wave = torch.istft(cplx, config.n_fft, hop_length=config.hop_length,
                       win_length=config.win_length, window=win, center=config.center,
                       length=length, normalized=False, onesided=True, return_complex=False)
    return wave.reshape(*lead, wave.shape[-1])


# ----------------------------------------------------------------------------- a5 masks
def mask_apply_ref(spec: torch.Tensor, mask: torch.Tensor, mode: str, pre_tanh: bool = False):
    """Mask application on [...,F,T,2] spectra.

    mode 'real': Y = X * m[...,None]                 (src/model/unet.py:62, dnn.py:140)
    mode 'E'   : polar, with the reference's 1e-8s   (src/model/dcunet.py:136-155, dccrn.py:203-217)
    mode 'C'   : complex multiply                    (dcunet.py:156-157, dccrn.py:218-219)
    mode 'R'   : re*mr, im*mi                        (dcunet.py:158-159, dccrn.py:220-221)
    pre_tanh   : DCUnet squashes the raw mask first  (dcunet.py:131)
    """
    xr, xi = spec[..., 0], spec[..., 1]
    if mode == "real":
        if pre_tanh:
            mask = torch.tanh(mask)
        return spec * mask.unsqueeze(-1)
    if pre_tanh:
        mask = torch.tanh(mask)
    mr, mi = mask[..., 0], mask[..., 1]
    if mode == "E":
        x_mag = torch.sqrt(xr ** 2 + xi ** 2 + 1e-8)
        x_ph = torch.atan2(xi, xr)
        m_mag = (mr ** 2 + mi ** 2) ** 0.5
        m_ph = torch.atan2(mi / (m_mag + 1e-8), mr / (m_mag + 1e-8))
        e_mag = torch.tanh(m_mag) * x_mag
        e_ph = x_ph + m_ph
        yr, yi = e_mag * torch.cos(e_ph), e_mag * torch.sin(e_ph)
    elif mode == "C":
        yr, yi = xr * mr - xi * mi, xr * mi + xi * mr
    elif mode == "R":
        yr, yi = xr * mr, xi * mi
    else:
        raise ValueError(f"unknown masking mode {mode!r}")
    return torch.stack([yr, yi], dim=-1)


def magnitude_feature_ref(spec: torch.Tensor, kind: str):
    """NN input features, quirks kept.

    'power'      |re^2+im^2|        src/model/unet.py:40
    'magnitude'  sqrt(re^2+im^2)    src/model/dnn.py:98
    'amplitude'  |re^2-im^2| (sic)  src/model/dcunet.py:379, stft_rnn.py:119, mel_rnn.py:123
    """
    re, im = spec[..., 0], spec[..., 1]
    if kind == "power":
        return torch.abs(re ** 2 + im ** 2)
    if kind == "magnitude":
        return torch.sqrt(re ** 2 + im ** 2)
    if kind == "amplitude":
        return torch.abs(re ** 2 - im ** 2)
    if kind == "crn":
        return torch.sqrt(re ** 2 - im ** 2)            # src/model/crn.py:101 (NaN source, kept)
    raise ValueError(kind)


# ----------------------------------------------------------------------------- a3 / a4 DCCRN
def _periodic_window(win_type, win_len):
    if win_type in (None, "None"):
        return np.ones(win_len)
    from scipy.signal import get_window
    return get_window(win_type, win_len, fftbins=True)


def conv_bases(win_len, fft_len, win_type="hann", inverse=False):
    """src/model/dccrn.py:649-666: windowed truncated-DFT basis (or its pinv, transposed)."""
    window = _periodic_window(win_type, win_len)
    dft_rows = np.fft.rfft(np.eye(fft_len))[:win_len]          # [win_len, F]
    basis = np.concatenate([dft_rows.real, dft_rows.imag], axis=1).T   # [2F, win_len]
    if inverse:
        basis = np.linalg.pinv(basis).T
    basis = basis * window
    return (torch.from_numpy(basis[:, None, :].astype(np.float32)),
            torch.from_numpy(window[None, :, None].astype(np.float32)))


def conv_stft_ref(wave, win_len, win_inc, fft_len, win_type="hann", feature_type="complex"):
    """src/model/dccrn.py:687-701.  [B,N] or [B,1,N] -> [B,2F,T] (or mags, phase)."""
    if wave.dim() == 2:
        wave = wave.unsqueeze(1)
    weight, _ = conv_bases(win_len, fft_len, win_type)
    pad = win_len - win_inc
    out = F.conv1d(F.pad(wave, [pad, pad]), weight.to(wave.dtype), stride=win_inc)
    if feature_type == "complex":
        return out
    nf = fft_len // 2 + 1
    re, im = out[:, :nf], out[:, nf:]
    return torch.sqrt(re ** 2 + im ** 2), torch.atan2(im, re)


def conv_istft_ref(spec, win_len, win_inc, fft_len, win_type="hann", length=None, phase=None):
    """src/model/dccrn.py:723-747.  [B,2F,T] -> [B,1,length or natural]."""
    if phase is not None:
        spec = torch.cat([spec * torch.cos(phase), spec * torch.sin(phase)], 1)
    weight, window = conv_bases(win_len, fft_len, win_type, inverse=True)
    weight, window = weight.to(spec.dtype), window.to(spec.dtype)
    out = F.conv_transpose1d(spec, weight, stride=win_inc)
    wsq = window.repeat(1, 1, spec.size(-1)) ** 2
    eye = torch.eye(win_len, dtype=spec.dtype)[:, None, :]
    env = F.conv_transpose1d(wsq, eye, stride=win_inc)
    out = out / (env + 1e-8)
    pad = win_len - win_inc
    if length:
        return out[..., pad:][..., :length]
    return out[..., pad:-pad]


# ----------------------------------------------------------------------------- MR-STFT loss
def mrstft_loss_ref(est: torch.Tensor, ref: torch.Tensor, resolutions=MRSTFT_RESOLUTIONS):
    """SURVEY.md 8(c) definition (the reference has no such loss; this is the pin).

    Per (n, hop, win): A = win * stft_custom(est), B = win * stft_custom(ref) (raw torch.stft
    with the reference's window / centre / reflect choices, src/evaluate.py:109-119);
    a = sqrt(clamp(|A|^2, 1e-7)); L_sc = ||b-a||_F / ||b||_F over the whole batch tensor;
    L_mag = mean |log b - log a|; loss = mean over resolutions of (L_sc + L_mag).
    Calling convention loss_function(enhanced, sources) -> 0-dim (src/solver.py:480).
    """
    total = est.new_zeros(())
    for n_fft, hop, win in resolutions:
        cfg = make_config(n_fft, hop, win)
        sa = stft_custom_ref(est, cfg) * win
        sb = stft_custom_ref(ref, cfg) * win
        a = torch.sqrt(torch.clamp(sa[..., 0] ** 2 + sa[..., 1] ** 2, min=MRSTFT_CLAMP))
        b = torch.sqrt(torch.clamp(sb[..., 0] ** 2 + sb[..., 1] ** 2, min=MRSTFT_CLAMP))
        l_sc = torch.linalg.norm((b - a).reshape(-1)) / torch.linalg.norm(b.reshape(-1))
        l_mag = torch.mean(torch.abs(torch.log(b) - torch.log(a)))
        total = total + l_sc + l_mag
    return total / len(resolutions)


def mrstft_partials_ref(est, ref, resolutions=MRSTFT_RESOLUTIONS):
    """The 3 partial sums per resolution that ranks exchange, float64."""
    out = []
    for n_fft, hop, win in resolutions:
        cfg = make_config(n_fft, hop, win)
        sa = (stft_custom_ref(est, cfg) * win).double()
        sb = (stft_custom_ref(ref, cfg) * win).double()
        a = torch.sqrt(torch.clamp(sa[..., 0] ** 2 + sa[..., 1] ** 2, min=MRSTFT_CLAMP))
        b = torch.sqrt(torch.clamp(sb[..., 0] ** 2 + sb[..., 1] ** 2, min=MRSTFT_CLAMP))
        out.append([float(((b - a) ** 2).sum()), float((b ** 2).sum()),
                    float(torch.abs(torch.log(b) - torch.log(a)).sum()), float(a.numel())])
    return out


# ----------------------------------------------------------------------------- evaluate() (8f-1)
def segment_ref(wave: torch.Tensor, num_feature: int, stride: int) -> torch.Tensor:
    """src/evaluate.py:164-183: zero-fill to a whole number of strides, cut overlapping segments."""
    n = wave.shape[-1]
    if n < num_feature:
        raise AssertionError("clip shorter than one segment")
    rem = (n - num_feature) % stride
    if rem:
        wave = F.pad(wave, [0, stride - rem])
    nseg = (wave.shape[-1] - num_feature) // stride + 1
    return torch.stack([wave[..., i * stride: i * stride + num_feature] for i in range(nseg)], 0)


def stitch_ref(segments: torch.Tensor, num_feature: int, stride: int, out_len: int):
    """src/evaluate.py:84-90: first segment whole, then the last `stride` samples of each later one."""
    nseg = segments.shape[0]
    out = torch.zeros(*segments.shape[1:-1], num_feature + stride * (nseg - 1), dtype=segments.dtype)
    out[..., :num_feature] = segments[0]
    for i in range(1, nseg):
        at = num_feature + stride * (i - 1)
        out[..., at: at + stride] = segments[i][..., -stride:]
    return out[..., :out_len]


def evaluate_ref(mixture, model, config, stft_models=("mel-rnn", "dcunet", "crn", "dnn", "unet", "rnn-stft-mask"),
                 monarch=("mel-rnn", "dcunet", "crn", "dnn", "unet", "dccrn", "wav-unet")):
    """src/evaluate.py:10-98 on CPU (single-source models): z-score, segment (stride = win_length),
    STFT, model on two half-batches, iSTFT, stitch, de-normalise."""
    with torch.no_grad():
        x = mixture
        if config.dset.norm == "z-score":
            mean = torch.mean(x, dim=-1, keepdim=True)
            std = torch.std(x, dim=-1, keepdim=True)
            x = (x - mean) / (std + 1e-9)
        stride = config.model.win_length
        nfeat = int(config.dset.sample_rate * config.model.segment)
        seg = segment_ref(x, nfeat, stride)
        nseg, nb, nc, ns = seg.shape
        batch = seg.reshape(nseg * nb, nc, ns)
        if config.model.name in stft_models:
            batch = stft_custom_ref(batch, config.model)
        if model:
            half = batch.shape[0] // 2
            out = torch.cat([model(batch[:half]), model(batch[half:])], 0)
        else:
            out = batch
        if config.model.name in monarch:
            out = out.unsqueeze(1)
        if config.model.name in stft_models:
            out = istft_custom_ref(out, ns, config.model)
        out = out.reshape(nseg, nb, nc, ns)
        enhanced = stitch_ref(out, nfeat, stride, mixture.shape[-1])
        if config.dset.norm == "z-score":
            enhanced = enhanced * (std + 1e-9) + mean
    return enhanced


def audio_seconds(rows_shape, sample_rate):
    """audio-seconds = clips * N / sample_rate (channels do not multiply)."""
    nbatch, nsample = rows_shape[0], rows_shape[-1]
    return nbatch * nsample / float(sample_rate)


def collate_fn_pad_ref(batch, segment_length, drop_last=True):
    """src/distrib.py:38-98 (+ pad_last, src/utils.py:12-15): pad short clips to one segment, drop or
    zero-pad the remainder, cut into segments, concatenate over the batch."""
    mixes, srcs, index_batch = [], [], []
    for item in batch:
        mixture, sources = item[0], item[1]
        if mixture.shape[-1] < segment_length:
            mixture = F.pad(mixture, [0, segment_length - mixture.shape[-1]])
            sources = F.pad(sources, [0, segment_length - sources.shape[-1]])
        rem = mixture.shape[-1] % segment_length
        if rem and drop_last:
            keep = segment_length * (mixture.shape[-1] // segment_length)
            mixture, sources = mixture[..., :keep], sources[..., :keep]
        elif rem:
            mixture = F.pad(mixture, [0, segment_length - rem])
            sources = F.pad(sources, [0, segment_length - rem])
        nch, length = mixture.shape
        nseg = length // segment_length
        mixes.append(mixture.reshape(nch, nseg, segment_length))
        srcs.append(sources.reshape(sources.shape[0], nch, nseg, segment_length))
        index_batch.append(nseg)
    return torch.cat(mixes, 1).permute(1, 0, 2), torch.cat(srcs, 2).permute(2, 0, 1, 3), index_batch


def si_snr_ref(s1, s2, eps=1e-8):
    """src/loss.py:17-29."""
    dot = torch.sum(s1 * s2, -1, keepdim=True)
    s2s2 = torch.sum(s2 * s2, -1, keepdim=True)
    s_target = dot / (s2s2 + eps) * s2
    e_noise = s1 - s_target
    tn = torch.sum(s_target * s_target, -1, keepdim=True)
    nn_ = torch.sum(e_noise * e_noise, -1, keepdim=True)
    return torch.mean(10 * torch.log10(tn / (nn_ + eps) + eps))


def psa_loss_ref(enhance, target, mixture):
    """src/loss.py:32-56 (eps 1e-9; the angles are tanh of the tangent, as written there)."""
    eps = 1e-9
    a_mix = torch.tanh(mixture[..., 1] / (mixture[..., 0] + eps))
    a_tgt = torch.tanh(target[..., 1] / (target[..., 0] + eps))
    amp_e = torch.sqrt(enhance[..., 1] ** 2 + enhance[..., 0] ** 2)
    amp_t = torch.sqrt(target[..., 1] ** 2 + target[..., 0] ** 2)
    return torch.mean((amp_e - amp_t * torch.cos(a_tgt - a_mix)) ** 2)


def overlap_and_add_ref(signal, frame_step):
    """src/model/conv_tasnet.py:11-31 restated without the sub-frame trick: frame f is added at offset
    f*frame_step, frames in increasing order (the order the reference's index_add_ applies its rows)."""
    outer = tuple(signal.shape[:-2])
    frames, length = signal.shape[-2:]
    out = signal.new_zeros(*outer, frame_step * (frames - 1) + length)
    for f in range(frames):
        out[..., f * frame_step:f * frame_step + length] += signal[..., f, :]
    return out


def si_sdr_metric_ref(reference, estimation):
    """SI_SDR of src/metric.py:92-123 (numpy there; mean of the per-row energy ratios, then dB; eps = the
    input dtype's machine epsilon)."""
    import numpy as np
    r = reference.detach().cpu().numpy() if torch.is_tensor(reference) else np.asarray(reference)
    e = estimation.detach().cpu().numpy() if torch.is_tensor(estimation) else np.asarray(estimation)
    energy = np.sum(r ** 2, axis=-1, keepdims=True)
    eps = np.finfo(energy.dtype).eps
    scale = np.sum(e * r, axis=-1, keepdims=True) / (energy + eps)
    proj = scale * r
    noise = e - proj
    ratio = np.mean(np.sum(proj ** 2, axis=-1) / (np.sum(noise ** 2, axis=-1) + eps))
    return 10 * np.log10(ratio + eps)
