"""Drop-in for the reference's STFT helpers: same names, arguments and shapes as
`stft_custom` / `istft_custom` in /root/reference/src/evaluate.py:101-162.

    from speech_enhancement_pytorch_b200.evaluate import stft_custom, istft_custom

`config` is duck-typed: any object with n_fft, hop_length, win_length, center
(src/conf/config.yaml:38-41, built by src/utils.py:149-165).
"""
from __future__ import annotations

from . import ops


def _cfg(config):
    if not getattr(config, "center", True):
        # the reference's own path always raises here: Hann + center=False has a zero envelope
        raise NotImplementedError("center=False is not built (torch.istft raises for it with a Hann window)")
    return int(config.n_fft), int(config.hop_length), int(config.win_length)


def stft_custom(tensor, config):
    """[B,C,N] or [B,S,C,N] -> [B,(S,)C,F,T,2], spectrum divided by win_length (evaluate.py:120)."""
    if tensor.dim() not in (3, 4):
        raise ValueError(f"stft_custom expects a 3-D or 4-D tensor, got {tensor.dim()}-D")
    n_fft, hop, win = _cfg(config)
    lead, nsample = tuple(tensor.shape[:-1]), tensor.shape[-1]
    spec = ops.stft(tensor.reshape(-1, nsample), n_fft, hop, win, 1.0 / win)
    return spec.reshape(*lead, *spec.shape[1:])


def istft_custom(tensor, length, config):
    """[B,C,F,T,2] or [B,S,C,F,T,2] -> [B,(S,)C,length] (evaluate.py:130-162)."""
    if tensor.dim() not in (5, 6):
        raise ValueError(f"istft_custom expects a 5-D or 6-D tensor, got {tensor.dim()}-D")
    n_fft, hop, win = _cfg(config)
    if tensor.shape[-1] != 2 or tensor.shape[-3] != n_fft // 2 + 1:
        raise RuntimeError(f"istft_custom: expected [..., {n_fft // 2 + 1}, T, 2], got {tuple(tensor.shape)}")
    lead = tuple(tensor.shape[:-3])
    nf, nt = tensor.shape[-3], tensor.shape[-2]
    if length is None:
        length = hop * (nt - 1)
    wave = ops.istft(tensor.reshape(-1, nf, nt, 2), int(length), n_fft, hop, win, float(win))
    return wave.reshape(*lead, wave.shape[-1])
