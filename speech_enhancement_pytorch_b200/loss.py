"""Multi-resolution STFT loss with the reference's loss calling convention
`loss_function(enhanced, sources) -> 0-dim tensor` (src/solver.py:480, src/loss.py:14-15;
selected in src/distrib.py:263-275 -- an `'mrstft'` key would return `loss_mrstft`).

The reference has no such loss; the definition is SURVEY.md 8(c):
spectral convergence + log-magnitude L1 over FFT sizes 512/1024/2048, hop n/4, Hann, reflect
centring, magnitudes clamped at sqrt(1e-7), Frobenius norms over the whole batch tensor.
"""
from __future__ import annotations

import torch

from . import ops


def loss_mrstft(enhanced, sources, group=None, global_rows=None, scale_grad_by_world=False):
    """enhanced, sources: waveforms [B,(S,)C,N] (any leading dims).  Gradient flows to `enhanced`.

    group: optional torch.distributed process group; when given, the partial sums AND the row counts of all ranks are
    summed in the path's only exchange step, so every rank gets the batch-global loss with the true denominator even
    when the ranks hold different numbers of rows (`global_rows` is accepted for compatibility and ignored).

    Gradient scale under data parallelism: each rank's gradient is d(GLOBAL loss)/d(its local rows).  Summed over
    ranks that is the single-GPU gradient.  DistributedDataParallel AVERAGES parameter gradients over ranks, so a
    model trained with DDP sees 1/world of the single-GPU gradient; pass scale_grad_by_world=True (multiplies the
    backward by the world size, leaves the loss value alone) to get the single-GPU gradient after DDP's averaging.
    """
    if enhanced.shape != sources.shape:
        raise ValueError(f"shape mismatch {tuple(enhanced.shape)} vs {tuple(sources.shape)}")
    n = enhanced.shape[-1]
    return ops.mrstft_loss_rows(enhanced.reshape(-1, n), sources.reshape(-1, n), group, scale_grad_by_world)


class MRSTFTLoss(torch.nn.Module):
    def __init__(self, group=None, scale_grad_by_world=False):
        super().__init__()
        self.group = group
        self.scale_grad_by_world = scale_grad_by_world

    def forward(self, enhanced, sources):
        return loss_mrstft(enhanced, sources, self.group, scale_grad_by_world=self.scale_grad_by_world)


def loss_spectral(enhanced_spec, sources_wave, config, kind="mse", group=None):
    """`loss_function(enhanced, stft_custom(sources, config))` for the reference's STFT-domain training
    losses (`mse` / `l1`, src/distrib.py:263-267 applied at src/solver.py:457-458,480) with the target
    spectrum never written to memory.

    enhanced_spec [B,(S,)C,F,T,2]; sources_wave [B,(S,)C,N] -> 0-dim tensor, gradient to enhanced_spec."""
    from .evaluate import _cfg
    kinds = {"mse": 0, "l1": 1}
    if kind not in kinds:
        raise ValueError(f"unknown spectral loss {kind!r}")
    n_fft, hop, win = _cfg(config)
    n = sources_wave.shape[-1]
    nf, nt = n_fft // 2 + 1, 1 + n // hop
    if tuple(enhanced_spec.shape[-3:]) != (nf, nt, 2) or enhanced_spec.shape[:-3] != sources_wave.shape[:-1]:
        raise ValueError(f"spectrum {tuple(enhanced_spec.shape)} does not match waveform {tuple(sources_wave.shape)}")
    return ops.spectral_loss_rows(enhanced_spec.reshape(-1, nf, nt, 2), sources_wave.reshape(-1, n), n_fft, hop, win,
                                  kinds[kind], group)


def si_snr(s1, s2, eps=1e-8):
    """`si_snr(s1, s2, eps)` of src/loss.py:21-29 (mean scale-invariant SNR in dB over all rows) in one pass
    over the two waveforms.  s1, s2 [...,N]."""
    if s1.shape != s2.shape:
        raise ValueError(f"shape mismatch {tuple(s1.shape)} vs {tuple(s2.shape)}")
    n = s1.shape[-1]
    return ops.si_snr_rows(s1.reshape(-1, n), s2.reshape(-1, n), eps)


def loss_sisdr(inputs, targets):
    """`loss_sisdr` of src/loss.py:14-15."""
    return -si_snr(inputs, targets)


def loss_phase_sensitive_spectral_approximation(enhance, target, mixture, group=None):
    """`loss_phase_sensitive_spectral_approximation(enhance, target, mixture)` of src/loss.py:32-56 (the
    `psa` loss, src/distrib.py:271-272; called with the mixture as third argument, src/solver.py:477-480),
    one fused elementwise-and-reduce pass over the three spectra [...,F,T,2]."""
    return ops.psa_loss(enhance, target, mixture, group)


def SI_SDR(reference, estimation, sr=16000):
    """`SI_SDR(reference, estimation)` of src/metric.py:92-123 on device tensors [...,T]: the three inner products
    per row come from one pass over both waveforms (the SI-SNR kernel), the per-row ratio, its mean and the
    10 log10 follow the reference's order (mean of the energy RATIOS, then dB; eps = float32 machine epsilon).
    Returns a 0-dim float64 tensor on the device (no host sync); no gradient (a metric)."""
    import torch
    if reference.shape != estimation.shape:
        raise ValueError(f"shape mismatch {tuple(reference.shape)} vs {tuple(estimation.shape)}")
    n = reference.shape[-1]
    with torch.no_grad():
        d = ops.row_dots(estimation.reshape(-1, n), reference.reshape(-1, n))      # <e,e>, <e,r>, <r,r>
        eps = float(torch.finfo(torch.float32).eps)
        ee, er, rr = d[:, 0], d[:, 1], d[:, 2]
        alpha = er / (rr + eps)
        proj = alpha * alpha * rr
        noise = (ee - 2.0 * alpha * er + proj).clamp_min(0.0)
        ratio = (proj / (noise + eps)).mean()
        return 10.0 * torch.log10(ratio + eps)
