"""Fused wave -> STFT -> mask -> iSTFT -> wave: what `istft_custom(model_tail(stft_custom(x)))`
computes (src/evaluate.py:39,72 around a mask-producing model), in one kernel launch; the
spectrum is never written to HBM (SURVEY.md 8d "fused enhance": 2S+M bytes instead of 2S+4P+M).
"""
from __future__ import annotations

from . import ops
from .evaluate import _cfg


def enhance(mixture, mask, config, mode="E", pre_tanh=False):
    """mixture [B,(S,)C,N]; mask [B,(S,)C,F,T] ('real') or [B,(S,)C,F,T,2] -> [B,(S,)C,N]."""
    n_fft, hop, win = _cfg(config)
    n = mixture.shape[-1]
    nf, nt = n_fft // 2 + 1, 1 + n // hop
    tail = (nf, nt) if mode == "real" else (nf, nt, 2)
    if tuple(mask.shape[-len(tail):]) != tail:
        raise ValueError(f"mask tail {tuple(mask.shape)} does not match {tail}")
    y = ops.enhance_rows(mixture.reshape(-1, n), mask.reshape(-1, *tail), n_fft, hop, win, mode, pre_tanh)
    return y.reshape(mixture.shape)
