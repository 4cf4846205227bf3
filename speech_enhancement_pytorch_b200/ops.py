"""Row-level ops over the C-ABI, with autograd.  Shapes are already flattened to rows here;
the reference-facing signatures live in evaluate.py / masking.py / loss.py / dccrn.py.

The step's hot operators (stft / istft / mask / mask+istft tail / enhance / single-process MR-STFT loss) go through the
PyTorch C++ extension (csrc_torch/se_torch.cpp: TORCH_LIBRARY(se_b200) + C++ autograd nodes over the same C-ABI);
the remaining operators bind the C-ABI with ctypes and Python autograd.Functions.
"""
from __future__ import annotations

import torch

from . import _native as nv


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def is_tuned(n_fft, hop):
    """True where the tuned engine is compiled (n_fft 512/1024/2048 at hop n/4 or n/2); every other power-of-two
    geometry runs on the general path behind the same entry points (csrc/se_generic.cuh)."""
    return n_fft in (512, 1024, 2048) and (hop * 4 == n_fft or hop * 2 == n_fft)


def _check_cfg(n_fft, hop, win_length, tuned_only=None):
    """tuned_only: name of a fused op that only exists for the tuned geometries."""
    if n_fft < 8 or n_fft > 8192 or (n_fft & 1):
        raise NotImplementedError(f"n_fft={n_fft}: even sizes in 8..8192 are built (no CPU / cuFFT fallback path)")
    if not (1 <= hop <= n_fft):
        raise NotImplementedError(f"hop_length={hop} must be in [1, n_fft]")
    if not (2 <= win_length <= n_fft):
        raise NotImplementedError(f"win_length={win_length} must be in [2, n_fft]")
    if tuned_only and not is_tuned(n_fft, hop):
        raise NotImplementedError(f"{tuned_only}: built for n_fft 512/1024/2048 at hop n_fft/4 or n_fft/2 (got {n_fft}/{hop}); "
                                  "the plain transforms (stft_custom / istft_custom / apply_mask) take any geometry")


# ------------------------------------------------------------------ raw calls
def stft_rows(x, n_fft, hop, win_length, scale):
    nv.require_cuda_f32(x)
    rows, n = x.shape
    out = torch.empty((rows, n_fft // 2 + 1, 1 + n // hop, 2), dtype=torch.float32, device=x.device)
    with nv.on_device(x.device):
        nv.check(nv.lib().se_stft_fwd(x.data_ptr(), out.data_ptr(), rows, n, n_fft, hop, win_length, scale,
                                      nv.stream_ptr(x.device)))
    return out


def stft_rows_adjoint(g, nsample, n_fft, hop, win_length, scale, out=None):
    nv.require_cuda_f32(g, out)
    rows = g.shape[0]
    acc = out is not None
    if out is None:
        out = torch.empty((rows, nsample), dtype=torch.float32, device=g.device)
    with nv.on_device(g.device):
        nv.check(nv.lib().se_stft_bwd(g.data_ptr(), out.data_ptr(), rows, nsample, n_fft, hop, win_length, scale,
                                      int(acc), nv.stream_ptr(g.device)))
    return out


def istft_rows(spec, length, n_fft, hop, win_length, scale):
    nv.require_cuda_f32(spec)
    rows, nf, nt, _ = spec.shape
    out = torch.empty((rows, length), dtype=torch.float32, device=spec.device)
    with nv.on_device(spec.device):
        nv.check(nv.lib().se_istft_fwd(spec.data_ptr(), out.data_ptr(), rows, nt, length, n_fft, hop, win_length,
                                       scale, nv.stream_ptr(spec.device)))
    return out


def istft_rows_adjoint(gy, nframe, n_fft, hop, win_length, scale):
    nv.require_cuda_f32(gy)
    rows, length = gy.shape
    out = torch.empty((rows, n_fft // 2 + 1, nframe, 2), dtype=torch.float32, device=gy.device)
    with nv.on_device(gy.device):
        nv.check(nv.lib().se_istft_bwd(gy.data_ptr(), out.data_ptr(), rows, nframe, length, n_fft, hop, win_length,
                                       scale, nv.stream_ptr(gy.device)))
    return out


# ------------------------------------------------------------------ autograd
def _as_f32(t):
    if t.dtype in (torch.float16, torch.bfloat16):
        return t.float()          # spectra stay fp32 under a bf16/fp16 model (BASELINE config 4)
    return t


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t.requires_grad for t in tensors)


def stft(x_rows, n_fft, hop, win_length, scale):
    # C++ op: config check, fp16/bf16 up-cast, contiguity, autograd node only when a gradient is needed
    return nv.torch_ops().stft(x_rows, n_fft, hop, win_length, float(scale))


def stft_nocenter(x_rows, n_fft, hop, win_length, scale):
    """torch.stft(center=False): frame t = x[t hop : t hop + n_fft] (general-geometry kernels for every size)."""
    _check_cfg(n_fft, hop, win_length)
    if x_rows.shape[-1] < n_fft:
        raise RuntimeError(f"stft (center=False): the input ({x_rows.shape[-1]} samples) is shorter than n_fft = {n_fft}")
    return nv.torch_ops().stft_nocenter(x_rows, n_fft, hop, win_length, float(scale))


def istft(spec_rows, length, n_fft, hop, win_length, scale):
    return nv.torch_ops().istft(spec_rows, int(length), n_fft, hop, win_length, float(scale))


class _MaskPlanar(torch.autograd.Function):
    """DCCRN layout: specs [rows,2F,T] planar, masks [rows,F,T] x 2 -> out [rows,2F,T] (one launch each way)."""

    @staticmethod
    def forward(ctx, specs, mre, mim, mode):
        nv.require_cuda_f32(specs, mre, mim)
        rows, nf2, nt = specs.shape
        out = torch.empty_like(specs)
        with nv.on_device(specs.device):
            nv.check(nv.lib().se_mask_planar_fwd(specs.data_ptr(), mre.data_ptr(), mim.data_ptr(), out.data_ptr(), rows,
                                                 nf2 // 2, nt, mode, nv.stream_ptr(specs.device)))
        ctx.save_for_backward(specs, mre, mim)
        ctx.mode = mode
        return out

    @staticmethod
    def backward(ctx, g):
        specs, mre, mim = ctx.saved_tensors
        rows, nf2, nt = specs.shape
        g = g.contiguous()
        gre, gim = torch.empty_like(mre), torch.empty_like(mim)
        gspec = torch.empty_like(specs) if ctx.needs_input_grad[0] else None
        with nv.on_device(specs.device):
            nv.check(nv.lib().se_mask_planar_bwd(specs.data_ptr(), mre.data_ptr(), mim.data_ptr(), g.data_ptr(), gre.data_ptr(),
                                                 gim.data_ptr(), _ptr(gspec), rows, nf2 // 2, nt, ctx.mode,
                                                 nv.stream_ptr(specs.device)))
        return gspec, gre, gim, None


def mask_apply_planar(specs, mask_real, mask_imag, mode):
    """specs [..., 2F, T] (Re bins then Im bins), masks [..., F, T]; returns a tensor shaped like specs."""
    if mode not in ("E", "C", "R"):
        raise ValueError(f"unknown DCCRN masking mode {mode!r}")
    if specs.dim() < 2 or specs.shape[-2] % 2:
        raise ValueError(f"expected specs [..., 2F, T], got {tuple(specs.shape)}")
    want = tuple(specs.shape[:-2]) + (specs.shape[-2] // 2, specs.shape[-1])
    if tuple(mask_real.shape) != want or tuple(mask_imag.shape) != want:
        raise ValueError(f"mask shapes {tuple(mask_real.shape)}, {tuple(mask_imag.shape)} do not match specs {tuple(specs.shape)}")
    lead = specs.shape[:-2]
    s3 = _as_f32(specs).contiguous().reshape(-1, specs.shape[-2], specs.shape[-1])
    m3 = [_as_f32(m).contiguous().reshape(-1, want[-2], want[-1]) for m in (mask_real, mask_imag)]
    return _MaskPlanar.apply(s3, m3[0], m3[1], nv.MASK_MODES[mode]).reshape(*lead, specs.shape[-2], specs.shape[-1])


def mask_apply(spec, mask, mode, pre_tanh=False):
    """spec [...,2], mask [...] ('real') or [...,2]; returns a tensor shaped like spec."""
    if mode not in nv.MASK_MODES:
        raise ValueError(f"unknown masking mode {mode!r}")
    want = spec.shape[:-1] if mode == "real" else spec.shape
    if tuple(mask.shape) != tuple(want):
        raise ValueError(f"mask shape {tuple(mask.shape)} does not match spectrum {tuple(spec.shape)} for mode {mode}")
    return nv.torch_ops().mask(spec, mask, nv.MASK_MODES[mode], bool(pre_tanh))


class _MRSTFT(torch.autograd.Function):
    """Utterance-sharded MR-STFT loss (group given).  The exchange step carries 10 doubles: the 9 partial sums and this
    rank's row count, so the mean's denominator is the TRUE global row count even when the shards are uneven
    (n_utterances % world != 0) -- it never leaves the device (no host sync, no guess like rows * world)."""

    @staticmethod
    def forward(ctx, est, ref, group, grad_scale):
        nv.require_cuda_f32(est, ref)
        rows, n = est.shape
        L = nv.lib()
        import torch.distributed as dist
        from . import distributed as sed
        ws = torch.empty(max(int(L.se_mrstft_workspace_bytes(rows, n)), 8), dtype=torch.uint8, device=est.device)
        sums = torch.empty(10, dtype=torch.float64, device=est.device)
        sums[9] = float(rows)                                   # this rank's rows; the exchange leaves the global count
        loss = torch.empty((), dtype=torch.float32, device=est.device)
        with nv.on_device(est.device):
            st = nv.stream_ptr(est.device)
            nv.check(L.se_mrstft_loss_fwd(est.data_ptr(), ref.data_ptr(), rows, n, sums.data_ptr(), ws.data_ptr(), st))
            # the one exchange step of the path, no host sync
            px = sed.peer_exchange(group, est.device)
            if px is not None:
                px.exchange_value(sums, None, n, loss, st)      # exchange + value in one kernel over peer memory
            else:
                dist.all_reduce(sums, group=group)              # NCCL: multi-node groups, or no peer access
                nv.check(L.se_mrstft_loss_value_dev(sums.data_ptr(), n, loss.data_ptr(), st))
        ctx.save_for_backward(est, ws, sums)
        ctx.grad_scale = float(grad_scale)
        return loss

    @staticmethod
    def backward(ctx, gout):
        est, ws, sums = ctx.saved_tensors
        rows, n = est.shape
        gout = gout.contiguous().float()
        if ctx.grad_scale != 1.0:
            gout = gout * ctx.grad_scale
        g = torch.empty_like(est)
        with nv.on_device(est.device):
            # global_rows = 0: the kernels read the global row count from sums[9]
            nv.check(nv.lib().se_mrstft_loss_bwd(est.data_ptr(), ws.data_ptr(), sums.data_ptr(), gout.data_ptr(),
                                                 0, rows, n, g.data_ptr(), nv.stream_ptr(est.device)))
        return g, None, None, None


def mrstft_loss_rows(est_rows, ref_rows, group=None, scale_grad_by_world=False):
    if ref_rows.requires_grad:
        raise NotImplementedError("loss_mrstft: gradient flows to `enhanced` only (targets must not require grad)")
    if group is None:
        return nv.torch_ops().mrstft_loss(est_rows, ref_rows)
    import torch.distributed as dist
    scale = float(dist.get_world_size(group)) if scale_grad_by_world else 1.0
    return _MRSTFT.apply(_as_f32(est_rows).contiguous(), _as_f32(ref_rows).contiguous(), group, scale)


def _global_count(local_count, total, group):
    """Sum (total, local element count) over the group in one all-reduce; returns (total, count) as device doubles --
    uneven shards get the true denominator without a host sync."""
    import torch.distributed as dist
    both = torch.stack([total.reshape(()), torch.tensor(float(local_count), dtype=torch.float64, device=total.device)])
    dist.all_reduce(both, group=group)
    return both[0], both[1]


class _SpectralLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enh, target, n_fft, hop, win_length, kind, group):
        nv.require_cuda_f32(enh, target)
        rows, n = target.shape
        L = nv.lib()
        ws = torch.empty(max(int(L.se_spectral_loss_workspace_bytes(rows, n, hop)), 8), dtype=torch.uint8, device=enh.device)
        total = torch.empty((), dtype=torch.float64, device=enh.device)
        with nv.on_device(enh.device):
            nv.check(L.se_spectral_loss_fwd(enh.data_ptr(), target.data_ptr(), rows, n, n_fft, hop, win_length,
                                            1.0 / win_length, kind, total.data_ptr(), ws.data_ptr(),
                                            nv.stream_ptr(enh.device)))
        count = rows * (n_fft // 2 + 1) * (1 + n // hop) * 2
        ratio = None
        if group is not None:
            # mean over the GLOBAL element count, summed on the device together with the partial sums (uneven shards)
            total, gcount = _global_count(count, total, group)
            ratio = (count / gcount).float()                    # local / global: folded into the upstream gradient
            loss = (total / gcount).float()
        else:
            loss = (total / count).float()
        ctx.save_for_backward(enh, target, ratio)
        ctx.cfg = (n_fft, hop, win_length, kind)
        return loss

    @staticmethod
    def backward(ctx, gout):
        enh, target, ratio = ctx.saved_tensors
        n_fft, hop, win_length, kind = ctx.cfg
        rows, n = target.shape
        gout = gout.contiguous().float()
        if ratio is not None:
            gout = gout * ratio                                 # d mean / d enh = (1 / global count) d sum / d enh
        g = torch.empty_like(enh)
        with nv.on_device(enh.device):
            nv.check(nv.lib().se_spectral_loss_bwd(enh.data_ptr(), target.data_ptr(), gout.data_ptr(), rows, rows, n,
                                                   n_fft, hop, win_length, 1.0 / win_length, kind, g.data_ptr(),
                                                   nv.stream_ptr(enh.device)))
        return g, None, None, None, None, None, None


def spectral_loss_rows(enh_rows, target_rows, n_fft, hop, win_length, kind, group=None):
    _check_cfg(n_fft, hop, win_length, tuned_only="spectral loss")
    if target_rows.requires_grad:
        raise NotImplementedError("spectral loss: gradient flows to the enhanced spectrum only")
    return _SpectralLoss.apply(_as_f32(enh_rows).contiguous(), _as_f32(target_rows).contiguous(), n_fft, hop, win_length,
                               kind, group)


class _SiSnr(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s1, s2, eps):
        nv.require_cuda_f32(s1, s2)
        rows, n = s1.shape
        dots = torch.empty(rows, 3, dtype=torch.float64, device=s1.device)
        snr = torch.empty(rows, dtype=torch.float32, device=s1.device)
        with nv.on_device(s1.device):
            nv.check(nv.lib().se_sisnr_fwd(s1.data_ptr(), s2.data_ptr(), rows, n, float(eps), dots.data_ptr(), snr.data_ptr(),
                                           nv.stream_ptr(s1.device)))
        ctx.save_for_backward(s1, s2, dots)
        ctx.eps = float(eps)
        return snr.mean()

    @staticmethod
    def backward(ctx, gout):
        s1, s2, dots = ctx.saved_tensors
        rows, n = s1.shape
        g = torch.empty_like(s1)
        gout = gout.contiguous().float()
        with nv.on_device(s1.device):
            nv.check(nv.lib().se_sisnr_bwd(s1.data_ptr(), s2.data_ptr(), dots.data_ptr(), gout.data_ptr(), 1.0 / rows, rows, n,
                                           ctx.eps, g.data_ptr(), nv.stream_ptr(s1.device)))
        return g, None, None


def si_snr_rows(s1_rows, s2_rows, eps=1e-8):
    if s2_rows.requires_grad:
        raise NotImplementedError("si_snr: gradient flows to the first argument only")
    return _SiSnr.apply(_as_f32(s1_rows).contiguous(), _as_f32(s2_rows).contiguous(), eps)


class _PSA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enh, tgt, mix, group):
        nv.require_cuda_f32(enh, tgt, mix)
        count = enh.numel() // 2
        L = nv.lib()
        ws = torch.empty(max(int(L.se_psa_workspace_bytes(count)), 8), dtype=torch.uint8, device=enh.device)
        total = torch.empty((), dtype=torch.float64, device=enh.device)
        with nv.on_device(enh.device):
            nv.check(L.se_psa_loss_fwd(enh.data_ptr(), tgt.data_ptr(), mix.data_ptr(), count, total.data_ptr(), ws.data_ptr(),
                                       nv.stream_ptr(enh.device)))
        ratio = None
        if group is not None:
            total, gcount = _global_count(count, total, group)
            ratio = (count / gcount).float()
            loss = (total / gcount).float()
        else:
            loss = (total / count).float()
        ctx.save_for_backward(enh, tgt, mix, ratio)
        return loss

    @staticmethod
    def backward(ctx, gout):
        enh, tgt, mix, ratio = ctx.saved_tensors
        g = torch.empty_like(enh)
        gout = gout.contiguous().float()
        if ratio is not None:
            gout = gout * ratio
        with nv.on_device(enh.device):
            nv.check(nv.lib().se_psa_loss_bwd(enh.data_ptr(), tgt.data_ptr(), mix.data_ptr(), gout.data_ptr(), enh.numel() // 2,
                                              enh.numel() // 2, g.data_ptr(), nv.stream_ptr(enh.device)))
        return g, None, None, None


def psa_loss(enh, tgt, mix, group=None):
    if tgt.requires_grad or mix.requires_grad:
        raise NotImplementedError("psa loss: gradient flows to `enhance` only")
    if enh.shape != tgt.shape or enh.shape != mix.shape or enh.shape[-1] != 2:
        raise ValueError("psa loss expects three spectra of identical shape [...,2]")
    return _PSA.apply(_as_f32(enh).contiguous(), _as_f32(tgt).contiguous(), _as_f32(mix).contiguous(), group)


# ------------------------------------------------------------------ fused enhance
def enhance_rows(x_rows, mask_rows, n_fft, hop, win_length, mode, pre_tanh=False):
    _check_cfg(n_fft, hop, win_length)
    if mode not in nv.MASK_MODES:
        raise ValueError(f"unknown masking mode {mode!r}")
    if not is_tuned(n_fft, hop):
        # general geometry: the three stages run as three launches (each one differentiable)
        spec = stft(x_rows, n_fft, hop, win_length, 1.0 / win_length)
        return istft(mask_apply(spec, mask_rows, mode, pre_tanh), x_rows.shape[-1], n_fft, hop, win_length, float(win_length))
    return nv.torch_ops().enhance(x_rows, mask_rows, n_fft, hop, win_length, nv.MASK_MODES[mode], bool(pre_tanh))


def mask_istft_rows(spec, mask, length, n_fft, hop, win_length, scale, mode, pre_tanh=False):
    """spec [rows,F,T,2], mask [rows,F,T] ('real') or [rows,F,T,2] -> [rows,length]."""
    _check_cfg(n_fft, hop, win_length)
    if mode not in nv.MASK_MODES:
        raise ValueError(f"unknown masking mode {mode!r}")
    if spec.requires_grad or not is_tuned(n_fft, hop):
        # the spectrum itself is being trained through (or a general geometry): keep the two differentiable stages separate
        return istft(mask_apply(spec, mask, mode, pre_tanh), length, n_fft, hop, win_length, scale)
    return nv.torch_ops().mask_istft(spec, mask, int(length), n_fft, hop, win_length, float(scale), nv.MASK_MODES[mode],
                                     bool(pre_tanh))


class _OverlapAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, signal, frame_step):
        nv.require_cuda_f32(signal)
        rows, frames, length = signal.shape
        out = torch.empty((rows, frame_step * (frames - 1) + length), dtype=torch.float32, device=signal.device)
        with nv.on_device(signal.device):
            nv.check(nv.lib().se_overlap_add_fwd(signal.data_ptr(), out.data_ptr(), rows, frames, length, frame_step,
                                                 nv.stream_ptr(signal.device)))
        ctx.cfg = (rows, frames, length, frame_step)
        return out

    @staticmethod
    def backward(ctx, gout):
        rows, frames, length, frame_step = ctx.cfg
        gout = gout.contiguous()
        g = torch.empty((rows, frames, length), dtype=torch.float32, device=gout.device)
        with nv.on_device(gout.device):
            nv.check(nv.lib().se_overlap_add_bwd(gout.data_ptr(), g.data_ptr(), rows, frames, length, frame_step,
                                                 nv.stream_ptr(gout.device)))
        return g, None


def overlap_add_rows(signal_rows, frame_step):
    """signal [rows, frames, frame_length] -> [rows, frame_step*(frames-1) + frame_length]."""
    if int(frame_step) <= 0:
        raise ValueError("frame_step must be positive")
    return _OverlapAdd.apply(_as_f32(signal_rows).contiguous(), int(frame_step))


def row_dots(s1_rows, s2_rows):
    """<s1,s1>, <s1,s2>, <s2,s2> per row in float64 ([rows,3]) -- one pass over both waveforms (se_sisnr_fwd)."""
    s1, s2 = _as_f32(s1_rows).contiguous(), _as_f32(s2_rows).contiguous()
    nv.require_cuda_f32(s1, s2)
    rows, n = s1.shape
    dots = torch.empty(rows, 3, dtype=torch.float64, device=s1.device)
    snr = torch.empty(rows, dtype=torch.float32, device=s1.device)
    with nv.on_device(s1.device):
        nv.check(nv.lib().se_sisnr_fwd(s1.data_ptr(), s2.data_ptr(), rows, n, 1e-8, dots.data_ptr(), snr.data_ptr(),
                                       nv.stream_ptr(s1.device)))
    return dots


# ------------------------------------------------------------------ DCCRN conv transforms
def register_window(values):
    """Register window values (any scipy.signal.get_window type, computed on the host) with the library and return
    their id (> 0; identical values share one) for the DCCRN transforms' `window_id` argument."""
    import ctypes
    import numpy as np
    w = np.ascontiguousarray(np.asarray(values, dtype=np.float64))
    wid = nv.lib().se_register_window(w.ctypes.data_as(ctypes.c_void_p), int(w.shape[0]))
    if wid <= 0:
        nv.check(wid)
    return int(wid)


def conv_is_tuned(win_len, win_inc, fft_len):
    """DCCRN's 400/100/512 family has fused kernels; other geometries run on the general path."""
    return bool(nv.lib().se_conv_geometry_tuned(int(win_len), int(win_inc), int(fft_len)))


def conv_stft_rows(x, win_len, win_inc, fft_len, window_id=0):
    return nv.torch_ops().conv_stft(x, win_len, win_inc, fft_len, window_id)


def polar_from_planar(spec):
    """[rows, 2F, T] planar spectrum -> (mags, phase) [rows, F, T] in one launch (ConvSTFT feature_type='real')."""
    nv.require_cuda_f32(spec)
    rows, nf2, nt = spec.shape
    mags = torch.empty((rows, nf2 // 2, nt), dtype=torch.float32, device=spec.device)
    phase = torch.empty_like(mags)
    with nv.on_device(spec.device):
        nv.check(nv.lib().se_polar_from_planar(spec.data_ptr(), mags.data_ptr(), phase.data_ptr(), rows, nf2 // 2, nt,
                                               nv.stream_ptr(spec.device)))
    return mags, phase


class _PlanarFromPolar(torch.autograd.Function):
    """cat([mags cos(phase), mags sin(phase)], 1) (ConviSTFT(inputs, phase), dccrn.py:729-732) in one launch each way."""

    @staticmethod
    def forward(ctx, mags, phase):
        nv.require_cuda_f32(mags, phase)
        rows, nf, nt = mags.shape
        spec = torch.empty((rows, 2 * nf, nt), dtype=torch.float32, device=mags.device)
        with nv.on_device(mags.device):
            nv.check(nv.lib().se_planar_from_polar(mags.data_ptr(), phase.data_ptr(), spec.data_ptr(), rows, nf, nt,
                                                   nv.stream_ptr(mags.device)))
        ctx.save_for_backward(mags, phase)
        return spec

    @staticmethod
    def backward(ctx, g):
        mags, phase = ctx.saved_tensors
        rows, nf, nt = mags.shape
        g = g.contiguous()
        gm, gp = torch.empty_like(mags), torch.empty_like(phase)
        with nv.on_device(mags.device):
            nv.check(nv.lib().se_planar_from_polar_bwd(mags.data_ptr(), phase.data_ptr(), g.data_ptr(), gm.data_ptr(), gp.data_ptr(),
                                                       rows, nf, nt, nv.stream_ptr(mags.device)))
        return gm, gp


def planar_from_polar(mags, phase):
    if mags.shape != phase.shape or mags.dim() != 3:
        raise ValueError(f"expected mags and phase [B,F,T], got {tuple(mags.shape)} and {tuple(phase.shape)}")
    return _PlanarFromPolar.apply(_as_f32(mags).contiguous(), _as_f32(phase).contiguous())


def conv_mask_istft_rows(spec, mask_real, mask_imag, out_len, win_len, win_inc, fft_len, mode, window_id=0):
    return nv.torch_ops().conv_mask_istft(spec, mask_real, mask_imag, int(out_len), win_len, win_inc, fft_len, nv.MASK_MODES[mode],
                                          window_id)


def conv_istft_rows(spec, out_len, win_len, win_inc, fft_len, window_id=0):
    return nv.torch_ops().conv_istft(spec, int(out_len), win_len, win_inc, fft_len, window_id)
