"""Mask application as one fused op (the tails of the reference models' `forward`s):

  'real'  unet.py:62, dnn.py:140, stft_rnn.py:108-109, mel_rnn.py:113, crn.py:139-141
  'E'     dcunet.py:136-155, dccrn.py:203-217  (polar; 1e-8 placements kept)
  'C'     dcunet.py:156-157, dccrn.py:218-219
  'R'     dcunet.py:158-159, dccrn.py:220-221
  pre_tanh=True squashes the raw mask first (dcunet.py:131).
"""
from __future__ import annotations

import torch

from . import ops


def apply_mask(spec, mask, mode="E", pre_tanh=False):
    """spec [...,F,T,2]; mask [...,F,T] for 'real', [...,F,T,2] otherwise -> [...,F,T,2]."""
    return ops.mask_apply(spec, mask, mode, pre_tanh)


def apply_mask_dccrn(specs, mask_real, mask_imag, mode="E"):
    """DCCRN layout (dccrn.py:147-223): specs [B,2F,T] (Re bins then Im bins), masks [B,F,T];
    returns out_spec [B,2F,T].  One planar kernel each way (se_mask_planar_fwd/bwd): no stack / cat copies."""
    return ops.mask_apply_planar(specs, mask_real, mask_imag, mode)


def magnitude_feature(spec, kind):
    """NN input feature of a spectrum [...,F,T,2] -> [...,F,T], the reference's quirks kept:
    'power' |re^2+im^2| (unet.py:40), 'magnitude' sqrt(re^2+im^2) (dnn.py:98), 'amplitude' |re^2-im^2|
    (dcunet.py:379 `Amplitude`, stft_rnn.py:119, mel_rnn.py:123), 'crn' sqrt(re^2-im^2) (crn.py:101)."""
    from . import _native as nv
    if kind not in nv.FEATURE_KINDS:
        raise ValueError(f"unknown feature kind {kind!r}")
    if spec.requires_grad:
        raise NotImplementedError("magnitude_feature: the reference never differentiates the input spectrogram")
    s = ops._as_f32(spec).contiguous()
    nv.require_cuda_f32(s)
    out = torch.empty(s.shape[:-1], dtype=torch.float32, device=s.device)
    with nv.on_device(s.device):
        nv.check(nv.lib().se_magnitude_feature(s.data_ptr(), out.data_ptr(), out.numel(), nv.FEATURE_KINDS[kind],
                                               nv.stream_ptr(s.device)))
    return out
