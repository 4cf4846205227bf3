"""Conv-TasNet's decoder tail: `overlap_and_add(signal, frame_step)`,
src/model/conv_tasnet.py:11-31 (called at :203 with frame_step = L // 2)."""
from __future__ import annotations

from . import ops


def overlap_and_add(signal, frame_step):
    """signal [..., frames, frame_length] -> [..., frame_step*(frames-1) + frame_length]; differentiable."""
    if signal.dim() < 2:
        raise ValueError("overlap_and_add expects [..., frames, frame_length]")
    outer = tuple(signal.shape[:-2])
    frames, length = signal.shape[-2:]
    out = ops.overlap_add_rows(signal.reshape(-1, frames, length), frame_step)
    return out.reshape(*outer, out.shape[-1])
