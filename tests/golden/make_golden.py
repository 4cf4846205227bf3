#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference from /root/reference.

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

Only modules that import cleanly are touched: src.evaluate, src.model.dccrn,
src.model.dcunet, src.model.unet.  Inputs are seeded; inputs AND outputs are stored so the
fixtures do not depend on torch's RNG stream.  Mask fixtures are captured with forward hooks
from real DCCRN / DCUnet forwards (the mask expressions live inside `forward`,
src/model/dccrn.py:147-223, src/model/dcunet.py:131-161).
"""
import os
import sys
import warnings
from types import SimpleNamespace

import numpy as np
import torch

REF = os.environ.get("SE_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")
if not hasattr(np, "int"):          # src/model/dccrn.py:675 uses the removed np.int alias
    np.int = int

from src.evaluate import stft_custom, istft_custom, _prepare_input_wav_zero_filled  # noqa: E402
from src.model.dccrn import ConvSTFT, ConviSTFT, DCCRN  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cfg(n_fft, hop, win):
    return SimpleNamespace(n_fft=n_fft, hop_length=hop, win_length=win, center=True)


def save(name, **arrays):
    out = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})


def gen_stft_istft():
    g = torch.Generator().manual_seed(20261017)
    cases = [  # tag, shape, n_fft, hop, win, istft length
        ("n512", (2, 1, 2048 + 37), 512, 128, 512, None),
        ("n1024", (2, 1, 4096 + 255), 1024, 256, 1024, None),
        ("n2048", (1, 2, 8192 + 513), 2048, 512, 2048, None),
        ("n512_hop256", (2, 1, 3000), 512, 256, 512, None),
        ("n512_win400", (2, 1, 2500), 512, 128, 400, None),
        ("n512_4d", (2, 2, 1, 1999), 512, 128, 512, None),
        ("n512_short_len", (1, 1, 2048), 512, 128, 512, 1900),
        ("n512_long_len", (1, 1, 2048), 512, 128, 512, 2100),
        ("n1024_multiple", (1, 1, 4096), 1024, 256, 1024, None),
    ]
    for tag, shape, n, h, w, length in cases:
        x = torch.randn(*shape, generator=g)
        c = cfg(n, h, w)
        spec = stft_custom(x, c)
        length = shape[-1] if length is None else length
        y = istft_custom(spec, length, c)
        # a perturbed (non-consistent) spectrum exercises the true inverse, not just round trip
        spec2 = spec + 0.01 * torch.randn(spec.shape, generator=g)
        y2 = istft_custom(spec2, length, c)
        save(f"stft_{tag}", x=x, spec=spec, y=y, spec2=spec2, y2=y2,
             meta=np.array([n, h, w, length]))
    # structured signals: impulses at the reflect boundaries, zeros, DC, Nyquist
    n, h, w = 512, 128, 512
    N = 1536
    sig = torch.zeros(6, 1, N)
    sig[0, 0, 0] = 1.0
    sig[1, 0, N - 1] = 1.0
    sig[3, 0] = 1.0
    sig[4, 0] = torch.tensor([1.0, -1.0]).repeat(N // 2)
    sig[5, 0, 255] = 1.0
    c = cfg(n, h, w)
    spec = stft_custom(sig, c)
    save("stft_structured", x=sig, spec=spec, y=istft_custom(spec, N, c), meta=np.array([n, h, w, N]))


def gen_grads():
    """autograd through the reference helpers: the adjoints of SURVEY a8."""
    g = torch.Generator().manual_seed(7)
    for tag, N, n, h, w in (("n512", 1200, 512, 128, 512), ("n1024", 2300, 1024, 256, 1024),
                            ("n512_win400", 1100, 512, 128, 400)):
        c = cfg(n, h, w)
        x = torch.randn(2, 1, N, generator=g, requires_grad=True)
        spec = stft_custom(x, c)
        gspec = torch.randn(spec.shape, generator=g)
        (gx,) = torch.autograd.grad(spec, x, gspec)
        s = (spec.detach() + 0.01 * torch.randn(spec.shape, generator=g)).requires_grad_(True)
        y = istft_custom(s, N, c)
        gy = torch.randn(y.shape, generator=g)
        (gs,) = torch.autograd.grad(y, s, gy)
        save(f"grad_{tag}", x=x, gspec=gspec, gx=gx, s=s, gy=gy, gs=gs, meta=np.array([n, h, w, N]))


def gen_conv():
    g = torch.Generator().manual_seed(11)
    for tag, N, length in (("a", 1600, None), ("b", 1637, 1637), ("c", 3200, 3000)):
        x = torch.randn(2, 1, N, generator=g)
        st = ConvSTFT(400, 100, 512, "hann", "complex")
        ist = ConviSTFT(400, 100, 512, length, "hann", "complex")
        spec = st(x)
        spec2 = spec + 0.05 * torch.randn(spec.shape, generator=g)
        save(f"conv_{tag}", x=x, spec=spec, y=ist(spec), spec2=spec2, y2=ist(spec2),
             meta=np.array([400, 100, 512, -1 if length is None else length]))
    st = ConvSTFT(400, 100, 512, "hann", "real")
    x = torch.randn(1, 1, 1000, generator=g)
    mags, phase = st(x)
    ist = ConviSTFT(400, 100, 512, None, "hann", "real")
    save("conv_polar", x=x, mags=mags, phase=phase, y=ist(mags, phase), meta=np.array([400, 100, 512, -1]))


def gen_dccrn_masks():
    """Capture (specs, mask, masked spec, wav) from real DCCRN forwards for E / C / R."""
    torch.manual_seed(3)
    for mode in ("E", "C", "R"):
        net = DCCRN(rnn_units=32, masking_mode=mode, use_clstm=True, kernel_num=[8, 8, 16, 16, 32, 32],
                    win_len=400, win_inc=100, fft_len=512, win_type="hann", length=1600)
        net.eval()
        grab = {}
        net.stft.register_forward_hook(lambda m, i, o: grab.__setitem__("specs", o.detach()))
        net.decoder[-1].register_forward_hook(lambda m, i, o: grab.__setitem__("dec", o.detach()))
        net.istft.register_forward_pre_hook(lambda m, i: grab.__setitem__("out_spec", i[0].detach()))
        x = 0.3 * torch.randn(2, 1, 1600)
        with torch.no_grad():
            wav = net(x)
        dec = grab["dec"][..., 1:]                                   # dccrn.py:194
        mask = torch.nn.functional.pad(dec, [0, 0, 1, 0])            # dccrn.py:198-201 (DC bin zero)
        save(f"dccrn_mask_{mode}", x=x, specs=grab["specs"], mask_re=mask[:, 0], mask_im=mask[:, 1],
             out_spec=grab["out_spec"], wav=wav)


def gen_dcunet_mask():
    from src.model.dcunet import DCUnet
    import inspect
    torch.manual_seed(5)
    sig = inspect.signature(DCUnet.__init__)
    kwargs = {}
    for k, v in dict(input_type="complex", complex=True, model_complexity=45, model_depth=10,
                     data_type=True, padding_mode="zeros", masking_mode="E", sources=["clean"],
                     audio_channels=1).items():
        if k in sig.parameters:
            kwargs[k] = v
    try:
        net = DCUnet(**kwargs)
    except Exception as e:  # constructor signature differs; record why and skip
        print("DCUnet construct failed:", repr(e))
        return
    net.eval()
    grab = {}
    net.linear.register_forward_hook(lambda m, i, o: grab.__setitem__("raw", o.detach()))
    c = cfg(1024, 256, 1024)
    x = torch.randn(1, 1, 256 * 64)
    spec = stft_custom(x, c)
    with torch.no_grad():
        out = net(spec)
    raw = grab["raw"].transpose(2, 3)                                # dcunet.py:132
    save("dcunet_mask_E", spec=spec, raw_mask=raw, out=out)


def gen_segments():
    g = torch.Generator().manual_seed(13)
    wav = torch.randn(1, 2, 5000, generator=g)
    seg = _prepare_input_wav_zero_filled(wav, 2048, 512)
    save("segments", wav=wav, seg=seg, meta=np.array([2048, 512]))


def gen_tasnet_and_metric():
    """Conv-TasNet's overlap_and_add (src/model/conv_tasnet.py:11-31) and the SI_SDR metric (src/metric.py:92-123).
    src.metric imports pesq / pypesq / pystoi / museval at module level (absent here, unused by SI_SDR): they
    are stubbed for the import only."""
    import types
    from src.model.conv_tasnet import overlap_and_add
    for name in ("pesq", "pypesq", "pystoi", "museval", "museval.metrics"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["pesq"].pesq = sys.modules["pesq"].cypesq = None
    sys.modules["pypesq"].pesq = None
    sys.modules["pystoi"].stoi = None
    sys.modules["museval.metrics"].bss_eval = None
    sys.modules["museval"].metrics = sys.modules["museval.metrics"]
    from src.metric import SI_SDR
    g = torch.Generator().manual_seed(77)
    out = {}
    for tag, (shape, step) in {"half": ((2, 3, 50, 40), 20), "quarter": ((3, 37, 32), 8), "coprime": ((2, 9, 10), 4),
                               "gap": ((2, 6, 8), 12), "abut": ((1, 5, 16), 16)}.items():
        sig = torch.randn(*shape, generator=g)
        out[f"sig_{tag}"] = sig
        out[f"step_{tag}"] = np.array(step)
        out[f"out_{tag}"] = overlap_and_add(sig, step)
    ref = torch.randn(3, 2, 4000, generator=g)
    est = ref + 0.3 * torch.randn(3, 2, 4000, generator=g)
    out["sdr_ref"], out["sdr_est"] = ref, est
    out["sdr"] = np.array(SI_SDR(ref, est), dtype=np.float64)
    out["sdr_scaled"] = np.array(SI_SDR(ref, 0.01 * est + 0.5), dtype=np.float64)
    save("tasnet_metric", **out)


if __name__ == "__main__":
    gen_tasnet_and_metric()
    if "--only-new" in sys.argv:
        sys.exit(0)
    gen_stft_istft()
    gen_grads()
    gen_conv()
    gen_dccrn_masks()
    gen_dcunet_mask()
    gen_segments()
