"""Drop-in modules for DCCRN's in-model transforms, same constructor arguments, call signatures and registered buffers
as `ConvSTFT` / `ConviSTFT` in /root/reference/src/model/dccrn.py:669-747, computed with FFT kernels instead of a dense
[2F x win_len] convolution (36x fewer flops).

Swap them into a constructed reference DCCRN with:  model.stft = ConvSTFT(...);  model.istft = ConviSTFT(...).
Like the reference they register `weight` (and `window`, `enframe` for ConviSTFT) buffers (dccrn.py:681,714,720-721),
built the way `init_kernels` (:649-666) builds them, so a reference checkpoint loads with strict=True.  The buffers are
NOT what the kernels compute with: the window values go to the library once (any scipy.signal.get_window type, including
the constructors' default 'hamming'), the Fourier basis is the FFT.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops


def _fft_len(win_len, fft_len):
    return int(2 ** math.ceil(math.log2(win_len))) if fft_len is None else int(fft_len)


def _window(win_type, win_len):
    """The reference's window (dccrn.py:651-654): ones for None / 'None', else scipy.signal.get_window(..., fftbins=True)."""
    if win_type == "None" or win_type is None:
        return np.ones(win_len)
    from scipy.signal import get_window
    return np.asarray(get_window(win_type, win_len, fftbins=True), dtype=np.float64)


def _init_kernels(win_len, fft_len, window, invers=False):
    """`init_kernels` of the reference (dccrn.py:649-666): [2F, 1, win_len] conv weight and [1, win_len, 1] window."""
    basis = np.fft.rfft(np.eye(fft_len))[:win_len]
    kernel = np.concatenate([np.real(basis), np.imag(basis)], 1).T
    if invers:
        kernel = np.linalg.pinv(kernel).T
    kernel = (kernel * window)[:, None, :]
    return torch.from_numpy(kernel.astype(np.float32)), torch.from_numpy(window[None, :, None].astype(np.float32))


class ConvSTFT(torch.nn.Module):
    def __init__(self, win_len, win_inc, fft_len=None, win_type="hamming", feature_type="real", fix=True):
        super().__init__()
        self.fft_len = _fft_len(win_len, fft_len)
        window = _window(win_type, win_len)
        kernel, _ = _init_kernels(win_len, self.fft_len, window)
        self.register_buffer("weight", kernel)                 # state-dict compatibility only (see module docstring)
        self.feature_type = feature_type
        self.stride = win_inc
        self.win_len = win_len
        self.dim = self.fft_len
        self._window_values = window
        self._window_id = None

    def _wid(self):
        if self._window_id is None:
            self._window_id = ops.register_window(self._window_values)
        return self._window_id

    def forward(self, inputs):
        if inputs.dim() == 3:
            if inputs.shape[1] != 1:
                raise RuntimeError("ConvSTFT expects [B,N] or [B,1,N]")
            inputs = inputs[:, 0]
        # (the reference's conv1d is differentiable wrt the waveform; the op raises NotImplementedError for that)
        out = ops.conv_stft_rows(inputs, self.win_len, self.stride, self.fft_len, self._wid())
        if self.feature_type == "complex":
            return out
        return ops.polar_from_planar(out)                       # (mags, phase), dccrn.py:696-701, one launch


class ConviSTFT(torch.nn.Module):
    def __init__(self, win_len, win_inc, fft_len=None, length=None, win_type="hamming", feature_type="real", fix=True):
        super().__init__()
        self.fft_len = _fft_len(win_len, fft_len)
        self.length = length
        window = _window(win_type, win_len)
        kernel, win = _init_kernels(win_len, self.fft_len, window, invers=True)
        self.register_buffer("weight", kernel)                  # state-dict compatibility only
        self.feature_type = feature_type
        self.win_type = win_type
        self.win_len = win_len
        self.stride = win_inc
        self.dim = self.fft_len
        self.register_buffer("window", win)
        self.register_buffer("enframe", torch.eye(win_len)[:, None, :])
        self._window_values = window
        self._window_id = None

    def _wid(self):
        if self._window_id is None:
            self._window_id = ops.register_window(self._window_values)
        return self._window_id

    def forward(self, inputs, phase=None):
        if phase is not None:
            inputs = ops.planar_from_polar(inputs, phase)       # cat([mags cos, mags sin], 1), dccrn.py:729-732, one launch
        return ops.conv_istft_rows(inputs, self._out_len(inputs.shape[-1]), self.win_len, self.stride, self.fft_len,
                                   self._wid()).unsqueeze(1)

    def _out_len(self, nt):
        pad = self.win_len - self.stride
        natural = self.stride * (nt - 1) + self.win_len - 2 * pad
        return min(int(self.length), natural + pad) if self.length else natural      # dccrn.py:741-745 slices, never extends

    def forward_masked(self, specs, mask_real, mask_imag, mode="E"):
        """`self(apply_mask_dccrn(specs, mask_real, mask_imag, mode))` -- DCCRN's tail, src/model/dccrn.py:203-224 --
        in one launch each way: the masked spectrum (and its gradient) is never written.  specs [B,2F,T] is data
        (ConvSTFT of the mixture); when it requires grad the two-stage composition runs instead."""
        from .masking import apply_mask_dccrn
        if mode not in ("E", "C", "R"):
            raise ValueError(f"unknown DCCRN masking mode {mode!r}")
        if specs.requires_grad or not ops.conv_is_tuned(self.win_len, self.stride, self.fft_len):
            # (general geometries: mask kernel, then the transform)
            return self(apply_mask_dccrn(specs, mask_real, mask_imag, mode))
        want = (specs.shape[0], specs.shape[1] // 2, specs.shape[2])
        if specs.dim() != 3 or tuple(mask_real.shape) != want or tuple(mask_imag.shape) != want:
            raise ValueError(f"mask shapes {tuple(mask_real.shape)}, {tuple(mask_imag.shape)} do not match specs {tuple(specs.shape)}")
        return ops.conv_mask_istft_rows(specs, mask_real, mask_imag, self._out_len(specs.shape[-1]), self.win_len, self.stride,
                                        self.fft_len, mode, self._wid()).unsqueeze(1)
