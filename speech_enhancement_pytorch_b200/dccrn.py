"""Drop-in modules for DCCRN's in-model transforms, same constructor arguments and call
signatures as `ConvSTFT` / `ConviSTFT` in /root/reference/src/model/dccrn.py:669-747, computed
with FFT kernels instead of a dense [2F x win_len] convolution (36x fewer flops).

Swap them into a constructed reference DCCRN with:  model.stft = ConvSTFT(...);
model.istft = ConviSTFT(...).  The reference registers `weight` / `window` / `enframe` buffers
(dccrn.py:681,714,720-721); these modules hold no tensors, so load reference checkpoints with
strict=False (the reference's own `_preload_model` already does, src/solver.py:274-276).
"""
from __future__ import annotations

import math

import torch

from . import ops


def _fft_len(win_len, fft_len):
    return int(2 ** math.ceil(math.log2(win_len))) if fft_len is None else int(fft_len)


def _check_window(win_type):
    if win_type not in ("hann", "hanning"):
        raise NotImplementedError(f"win_type={win_type!r}: only the periodic Hann window used by DCCRN "
                                  "(dccrn.py:20) is built")


class ConvSTFT(torch.nn.Module):
    def __init__(self, win_len, win_inc, fft_len=None, win_type="hamming", feature_type="real", fix=True):
        super().__init__()
        _check_window(win_type)
        self.fft_len = _fft_len(win_len, fft_len)
        self.feature_type = feature_type
        self.stride = win_inc
        self.win_len = win_len
        self.dim = self.fft_len

    def forward(self, inputs):
        if inputs.dim() == 3:
            if inputs.shape[1] != 1:
                raise RuntimeError("ConvSTFT expects [B,N] or [B,1,N]")
            inputs = inputs[:, 0]
        out = ops.conv_stft_rows(ops._as_f32(inputs).contiguous(), self.win_len, self.stride, self.fft_len)
        if self.feature_type == "complex":
            return out
        nf = self.dim // 2 + 1
        real, imag = out[:, :nf, :], out[:, nf:, :]
        return torch.sqrt(real ** 2 + imag ** 2), torch.atan2(imag, real)


class ConviSTFT(torch.nn.Module):
    def __init__(self, win_len, win_inc, fft_len=None, length=None, win_type="hamming", feature_type="real", fix=True):
        super().__init__()
        _check_window(win_type)
        self.fft_len = _fft_len(win_len, fft_len)
        self.length = length
        self.feature_type = feature_type
        self.win_type = win_type
        self.win_len = win_len
        self.stride = win_inc
        self.dim = self.fft_len

    def forward(self, inputs, phase=None):
        if phase is not None:
            inputs = torch.cat([inputs * torch.cos(phase), inputs * torch.sin(phase)], 1)
        nt = inputs.shape[-1]
        pad = self.win_len - self.stride
        natural = self.stride * (nt - 1) + self.win_len - 2 * pad
        if self.length:
            out_len = min(int(self.length), natural + pad)     # dccrn.py:741-743 slices, never extends
        else:
            out_len = natural
        return ops.conv_istft_rows(inputs, out_len, self.win_len, self.stride, self.fft_len).unsqueeze(1)

    def _out_len(self, nt):
        pad = self.win_len - self.stride
        natural = self.stride * (nt - 1) + self.win_len - 2 * pad
        return min(int(self.length), natural + pad) if self.length else natural

    def forward_masked(self, specs, mask_real, mask_imag, mode="E"):
        """`self(apply_mask_dccrn(specs, mask_real, mask_imag, mode))` -- DCCRN's tail, src/model/dccrn.py:203-224 --
        in one launch each way: the masked spectrum (and its gradient) is never written.  specs [B,2F,T] is data
        (ConvSTFT of the mixture); when it requires grad the two-stage composition runs instead."""
        from .masking import apply_mask_dccrn
        if mode not in ("E", "C", "R"):
            raise ValueError(f"unknown DCCRN masking mode {mode!r}")
        if specs.requires_grad:
            return self(apply_mask_dccrn(specs, mask_real, mask_imag, mode))
        want = (specs.shape[0], specs.shape[1] // 2, specs.shape[2])
        if specs.dim() != 3 or tuple(mask_real.shape) != want or tuple(mask_imag.shape) != want:
            raise ValueError(f"mask shapes {tuple(mask_real.shape)}, {tuple(mask_imag.shape)} do not match specs {tuple(specs.shape)}")
        return ops.conv_mask_istft_rows(specs, mask_real, mask_imag, self._out_len(specs.shape[-1]), self.win_len, self.stride,
                                        self.fft_len, mode).unsqueeze(1)
