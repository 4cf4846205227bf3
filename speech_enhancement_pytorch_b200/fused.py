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


def apply_mask_istft(spec, mask, length, config, mode="E", pre_tanh=False):
    """`istft_custom(apply_mask(spec, mask, mode, pre_tanh), length, config)` -- the model tail followed by the
    inverse transform, as `evaluate()` (src/evaluate.py:54-72) and the training step (src/solver.py:466-480)
    chain them -- in one launch forward and one backward; the masked spectrum is never written.
    spec [B,(S,)C,F,T,2]; mask [B,(S,)C,F,T] ('real') or [...,F,T,2] -> [B,(S,)C,length]."""
    if spec.dim() not in (5, 6):
        raise ValueError(f"apply_mask_istft expects a 5-D or 6-D spectrum, got {spec.dim()}-D")
    n_fft, hop, win = _cfg(config)
    nf, nt = spec.shape[-3], spec.shape[-2]
    if spec.shape[-1] != 2 or nf != n_fft // 2 + 1:
        raise RuntimeError(f"apply_mask_istft: expected [..., {n_fft // 2 + 1}, T, 2], got {tuple(spec.shape)}")
    want = tuple(spec.shape[:-1]) if mode == "real" else tuple(spec.shape)
    if tuple(mask.shape) != want:
        raise ValueError(f"mask shape {tuple(mask.shape)} does not match spectrum {tuple(spec.shape)} for mode {mode}")
    if length is None:
        length = hop * (nt - 1)
    lead = tuple(spec.shape[:-3])
    tail = (nf, nt) if mode == "real" else (nf, nt, 2)
    y = ops.mask_istft_rows(spec.reshape(-1, nf, nt, 2), mask.reshape(-1, *tail), int(length), n_fft, hop, win,
                            float(win), mode, pre_tanh)
    return y.reshape(*lead, y.shape[-1])
