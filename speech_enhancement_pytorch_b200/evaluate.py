"""Drop-in for the reference's STFT helpers: same names, arguments and shapes as
`stft_custom` / `istft_custom` in /root/reference/src/evaluate.py:101-162.

    from speech_enhancement_pytorch_b200.evaluate import stft_custom, istft_custom

`config` is duck-typed: any object with n_fft, hop_length, win_length, center
(src/conf/config.yaml:38-41, built by src/utils.py:149-165).
"""
from __future__ import annotations

from . import ops


def _cfg(config, allow_nocenter=False):
    if not getattr(config, "center", True) and not allow_nocenter:
        # the reference's own path raises here too: torch.istft checks the overlap-add envelope, which is zero at the
        # first sample for a Hann window without centre padding (src/evaluate.py:143-152)
        raise RuntimeError("center=False: window overlap add min < 1e-11 (torch.istft raises the same for the reference's "
                           "Hann window); only stft_custom runs without centre padding")
    return int(config.n_fft), int(config.hop_length), int(config.win_length)


def stft_custom(tensor, config):
    """[B,C,N] or [B,S,C,N] -> [B,(S,)C,F,T,2], spectrum divided by win_length (evaluate.py:120)."""
    if tensor.dim() not in (3, 4):
        raise ValueError(f"stft_custom expects a 3-D or 4-D tensor, got {tensor.dim()}-D")
    n_fft, hop, win = _cfg(config, allow_nocenter=True)
    lead, nsample = tuple(tensor.shape[:-1]), tensor.shape[-1]
    if getattr(config, "center", True):
        spec = ops.stft(tensor.reshape(-1, nsample), n_fft, hop, win, 1.0 / win)
    else:                                                       # torch.stft(center=False): no padding (evaluate.py:116)
        spec = ops.stft_nocenter(tensor.reshape(-1, nsample), n_fft, hop, win, 1.0 / win)
    return spec.reshape(*lead, *spec.shape[1:])


def stft_custom_with_feature(tensor, config, kind):
    """`stft_custom` that also returns the model's magnitude feature [B,(S,)C,F,T],
    written by the same kernel while each bin is still in registers."""
    import torch
    from . import _native as nv
    if tensor.dim() not in (3, 4):
        raise ValueError(f"stft_custom expects a 3-D or 4-D tensor, got {tensor.dim()}-D")
    if kind not in nv.FEATURE_KINDS:
        raise ValueError(f"unknown feature kind {kind!r}")
    n_fft, hop, win = _cfg(config)
    ops._check_cfg(n_fft, hop, win)
    if torch.is_grad_enabled() and tensor.requires_grad:
        # torch.stft is differentiable wrt the waveform; this fused variant has no adjoint wired (use stft_custom)
        raise NotImplementedError("stft_custom_with_feature: gradient wrt the input waveform is not built")
    lead, nsample = tuple(tensor.shape[:-1]), tensor.shape[-1]
    x = ops._as_f32(tensor).reshape(-1, nsample).contiguous()
    nv.require_cuda_f32(x)
    rows = x.shape[0]
    nf, nt = n_fft // 2 + 1, 1 + nsample // hop
    if not ops.is_tuned(n_fft, hop):                            # general geometry: transform, then the feature kernel
        spec = ops.stft(x, n_fft, hop, win, 1.0 / win)
        feat = torch.empty((rows, nf, nt), dtype=torch.float32, device=x.device)
        with nv.on_device(x.device):
            nv.check(nv.lib().se_magnitude_feature(spec.data_ptr(), feat.data_ptr(), rows * nf * nt, nv.FEATURE_KINDS[kind],
                                                   nv.stream_ptr(x.device)))
        return spec.reshape(*lead, nf, nt, 2), feat.reshape(*lead, nf, nt)
    spec = torch.empty((rows, nf, nt, 2), dtype=torch.float32, device=x.device)
    feat = torch.empty((rows, nf, nt), dtype=torch.float32, device=x.device)
    with nv.on_device(x.device):
        nv.check(nv.lib().se_stft_feature_fwd(x.data_ptr(), spec.data_ptr(), feat.data_ptr(), rows, nsample, n_fft, hop,
                                              win, 1.0 / win, nv.FEATURE_KINDS[kind], nv.stream_ptr(x.device)))
    return spec.reshape(*lead, nf, nt, 2), feat.reshape(*lead, nf, nt)


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


# ------------------------------------------------------------------------------------------------
# evaluate(): the inference-side caller of the path, src/evaluate.py:10-98

MULTI_SPEECH_SEPERATION_MODELS = ("demucs", "conv-tasnet", "rnn-stft-mask")          # src/model/types.py:1
MONARCH_SPEECH_SEPARTAION_MODELS = ("mel-rnn", "dcunet", "crn", "dnn", "unet", "dccrn", "wav-unet")   # :3
STFT_MODELS = ("mel-rnn", "dcunet", "crn", "dnn", "unet", "rnn-stft-mask")           # src/model/types.py:5


def row_stats(wave):
    """Per-(batch, channel) mean and unbiased std of [B,C,L] in one launch: [B*C, 4] = (mean, 1/(std+1e-9), std+1e-9, 0)
    -- the z-score of src/evaluate.py:18-21, consumed by `segment_stft` / `istft_stitch` instead of being applied as
    separate elementwise passes."""
    import torch
    from . import _native as nv
    x = ops._as_f32(wave).contiguous()
    nv.require_cuda_f32(x)
    rows, length = x.shape[0] * x.shape[1], x.shape[2]
    stats = torch.empty((rows, 4), dtype=torch.float32, device=x.device)
    with nv.on_device(x.device):
        nv.check(nv.lib().se_row_stats(x.data_ptr(), stats.data_ptr(), rows, length, length, nv.stream_ptr(x.device)))
    return stats


def segment_stft(wave, num_feature, stride, config, stats=None):
    """`_prepare_input_wav_zero_filled` (src/evaluate.py:164-183) + reshape (:35-36) + `stft_custom` (:39)
    in one launch: the overlapping segments are never materialised (they are strided views of the clip;
    the zero-filled tail is synthesised in the kernel).  wave [B,C,L] -> [nseg*B, C, F, T, 2].
    stats (from `row_stats`): the z-score is applied to the samples while they are staged."""
    import torch
    from . import _native as nv
    n_fft, hop, win = _cfg(config)
    ops._check_cfg(n_fft, hop, win, tuned_only="evaluate() / segment_stft")
    if wave.dim() != 3:
        raise ValueError("segment_stft expects [B,C,L]")
    if torch.is_grad_enabled() and wave.requires_grad:
        raise NotImplementedError("segment_stft: gradient wrt the input waveform is not built (evaluate() runs under no_grad)")
    nb, nc, length = wave.shape
    if length < num_feature:
        raise AssertionError("the length of data is too short comparing the number of features...")
    rem = (length - num_feature) % stride
    padded = length + (stride - rem if rem else 0)
    nseg = (padded - num_feature) // stride + 1
    x = ops._as_f32(wave).contiguous()
    nv.require_cuda_f32(x, stats)
    nf, nt = n_fft // 2 + 1, 1 + num_feature // hop
    out = torch.empty((nseg * nb, nc, nf, nt, 2), dtype=torch.float32, device=x.device)
    L = nv.lib()
    sp = 0 if stats is None else stats.data_ptr()
    # Overlapping segments share all frames that do not touch their reflect padding: transform those ONCE at clip level and
    # gather (three launches), instead of transforming every frame of every segment (a 30 s clip at stride 512 has 814
    # segments x 501 frames of which 3 750 are distinct).  Falls back to the per-segment kernel when the stride is no
    # multiple of the hop or there is nothing to share.
    scratch_bytes = int(L.se_stft_segments_scratch_bytes(nseg, nb * nc, stride, num_feature, n_fft, hop)) if nseg > 2 else 0
    with nv.on_device(x.device):
        if scratch_bytes > 0:
            scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=x.device)
            nv.check(L.se_stft_segments_shared_fwd(x.data_ptr(), out.data_ptr(), sp, nb * nc, nb * nc, nseg, nb * nc, length, length,
                                                   stride, num_feature, n_fft, hop, win, 1.0 / win, scratch.data_ptr(),
                                                   nv.stream_ptr(x.device)))
        else:
            nv.check(L.se_stft_segments_norm_fwd(x.data_ptr(), out.data_ptr(), sp, nb * nc, nb * nc, nseg, nb * nc, length, length,
                                                 stride, num_feature, n_fft, hop, win, 1.0 / win, nv.stream_ptr(x.device)))
    return out, nseg


def istft_stitch(spec, nseg, num_feature, stride, out_len, config, stats=None, stats_channels=1):
    """istft_custom of every segment + the reference's stitch (src/evaluate.py:84-90: segment 0 whole, then the last `stride`
    samples of each later segment, trimmed to out_len) + de-normalisation (:92-93) in ONE launch that only synthesises the
    frames overlapping kept samples.  spec [nseg*K, ..., F, T, 2] (segment-major rows) -> [K*..., out_len].
    stats: `row_stats` of the [B,C,L] mixture; stats_channels = C (a sources dimension between B and C shares them)."""
    import torch
    from . import _native as nv
    n_fft, hop, win = _cfg(config)
    ops._check_cfg(n_fft, hop, win, tuned_only="evaluate() / istft_stitch")
    s = ops._as_f32(spec).contiguous()
    nv.require_cuda_f32(s, stats)
    nf, nt = s.shape[-3], s.shape[-2]
    rows = s.numel() // (nf * nt * 2)
    if s.shape[-1] != 2 or nf != n_fft // 2 + 1 or rows % nseg:
        raise RuntimeError(f"istft_stitch: expected [nseg*K, ..., {n_fft // 2 + 1}, T, 2] with nseg = {nseg}, got {tuple(spec.shape)}")
    nclip = rows // nseg
    out = torch.empty((nclip, out_len), dtype=torch.float32, device=s.device)
    div = 1 if stats is None else max(nclip * stats_channels // stats.shape[0], 1)      # = nsrc * C
    with nv.on_device(s.device):
        nv.check(nv.lib().se_istft_stitch_fwd(s.data_ptr(), out.data_ptr(), 0 if stats is None else stats.data_ptr(), div,
                                              stats_channels, nseg, nclip, nt, num_feature, stride, out_len, out_len, n_fft, hop,
                                              win, float(win), nv.stream_ptr(s.device)))
    return out


def stitch_segments(output, num_feature, stride, out_len):
    """src/evaluate.py:84-90 without the Python loop (waveform models; the STFT models stitch inside `istft_stitch`):
    the first segment whole, then the last `stride` samples of every later segment.  output [nseg, ..., num_feature]
    -> [..., out_len]."""
    import torch
    nseg = output.shape[0]
    head = output[0]
    if nseg == 1:
        return head[..., :out_len]
    tails = output[1:, ..., num_feature - stride:]                       # [nseg-1, ..., stride]
    tails = tails.movedim(0, -2).reshape(*output.shape[1:-1], (nseg - 1) * stride)
    return torch.cat([head, tails], dim=-1)[..., :out_len]


def evaluate(mixture, model, device, config):
    """Drop-in for `evaluate(mixture, model, device, config)` (src/evaluate.py:10-98): normalise,
    cut into `config.model.segment`-second segments with stride win_length, STFT, model, iSTFT,
    stitch, de-normalise.  Everything runs on `device` (the reference runs the STFT on the CPU tensor
    and moves to the device afterwards, :39,50).

    STFT models: three launches around the model -- `row_stats`, `segment_stft` (z-score folded into the staging of the
    strided segment views) and `istft_stitch` (inverse transform of only the frames the stitch keeps, de-normalised and
    written straight into the stitched clip).  The result comes back on `mixture`'s device (the reference builds it with
    torch.zeros on the CPU, :83; its test loop feeds CPU mixtures, src/solver.py:584)."""
    import torch
    with torch.no_grad():
        x = mixture.to(device)
        norm = getattr(config.dset, "norm", None)
        if norm == "linear-scale":
            # the reference indexes the (values, indices) tuple of torch.max here and fails (:23-25)
            raise NotImplementedError("dset.norm='linear-scale' is broken in the reference (src/evaluate.py:23-25)")
        stride = config.model.win_length
        num_feature = int(config.dset.sample_rate * config.model.segment)
        nbatch, nchannel, length = x.shape
        name = config.model.name
        multi = bool(model) and name in MULTI_SPEECH_SEPERATION_MODELS
        n_fft, hop, _ = _cfg(config.model) if name in STFT_MODELS else (0, 0, 0)
        if name in STFT_MODELS and ops.is_tuned(n_fft, hop):
            stats = row_stats(x) if norm == "z-score" else None
            batch, nseg = segment_stft(x, num_feature, stride, config.model, stats)
            if model:
                model.eval()
                half = int(batch.shape[0] // 2)                                  # :48-56 two half-batches
                output = torch.cat([model(batch[:half]), model(batch[half:])], dim=0)
            else:
                output = batch
            enhanced = istft_stitch(output, nseg, num_feature, stride, length, config.model, stats, nchannel)
            if multi:
                enhanced = enhanced.reshape(nbatch, len(config.model.sources), nchannel, length)
            else:
                enhanced = enhanced.reshape(nbatch, nchannel, length)
            return enhanced.to(mixture.device)
        # waveform models: no transform on the path -- normalise, view the segments, model, stitch.  STFT models at a
        # general geometry (no fused segment / stitch kernels) take the same route with the plain transforms around the model.
        spectral = name in STFT_MODELS
        if norm == "z-score":
            mean = torch.mean(x, dim=-1, keepdim=True)
            std = torch.std(x, dim=-1, keepdim=True)
            x = (x - mean) / (std + 1e-9)
        rem = (length - num_feature) % stride
        xp = torch.nn.functional.pad(x, [0, stride - rem]) if rem else x
        batch = xp.unfold(-1, num_feature, stride).movedim(-2, 0)             # [nseg,B,C,N] view
        nseg = batch.shape[0]
        batch = batch.reshape(nseg * nbatch, nchannel, num_feature)
        if spectral:
            batch = stft_custom(batch, config.model)
        if model:
            model.eval()
            half = int(batch.shape[0] // 2)
            output = torch.cat([model(batch[:half]), model(batch[half:])], dim=0)
        else:
            output = batch
        if spectral:
            output = istft_custom(output, num_feature, config.model)
        if name in MONARCH_SPEECH_SEPARTAION_MODELS:
            output = torch.unsqueeze(output, dim=1)
        if multi:
            output = output.reshape(nseg, nbatch, len(config.model.sources), nchannel, num_feature)
        else:
            output = output.reshape(nseg, nbatch, nchannel, num_feature)
        enhanced = stitch_segments(output, num_feature, stride, length)
        if norm == "z-score":
            if enhanced.dim() == mean.dim() + 1:
                mean, std = mean.unsqueeze(1), std.unsqueeze(1)
            enhanced = enhanced * (std + 1e-9) + mean
    return enhanced.to(mixture.device)
