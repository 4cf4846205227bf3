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


def gen_general_geometry():
    """The reference's helpers and DCCRN transforms at geometries its configs do not use (any n_fft / hop / win_length is
    accepted by torch.stft; any win_len / win_inc / fft_len by ConvSTFT): pins the general-geometry kernel path."""
    g = torch.Generator().manual_seed(4242)
    out = {}
    metas = []
    for i, (N, n, h, w) in enumerate(((1500, 256, 64, 256), (2400, 512, 160, 400), (1777, 512, 100, 512), (5000, 4096, 1024, 4096),
                                      (700, 64, 16, 64), (3000, 2048, 300, 1200), (3000, 320, 160, 320), (2222, 400, 100, 400))):
        c = cfg(n, h, w)
        x = torch.randn(2, 1, N, generator=g, requires_grad=True)
        spec = stft_custom(x, c)
        gspec = torch.randn(spec.shape, generator=g)
        (gx,) = torch.autograd.grad(spec, x, gspec)
        s = (spec.detach() + 0.01 * torch.randn(spec.shape, generator=g)).requires_grad_(True)
        y = istft_custom(s, N, c)
        gy = torch.randn(y.shape, generator=g)
        (gs,) = torch.autograd.grad(y, s, gy)
        for k, v in (("x", x), ("spec", spec), ("gspec", gspec), ("gx", gx), ("s", s), ("y", y), ("gy", gy), ("gs", gs)):
            out[f"t{i}_{k}"] = v
        metas.append([N, n, h, w])
    out["t_meta"] = np.array(metas)
    metas = []
    for i, (N, wl, inc, nfft, wt) in enumerate(((1600, 320, 160, 512, "hann"), (2000, 400, 100, 1024, "hamming"),
                                                (1200, 256, 64, 256, "hann"), (1700, 400, 128, 512, "hamming"))):
        x = torch.randn(2, 1, N, generator=g)
        st = ConvSTFT(wl, inc, nfft, wt, "complex")
        ist = ConviSTFT(wl, inc, nfft, None, wt, "complex")
        spec = st(x)
        s = (spec + 0.05 * torch.randn(spec.shape, generator=g)).detach().requires_grad_(True)
        y = ist(s)
        gy = torch.randn(y.shape, generator=g)
        (gs,) = torch.autograd.grad(y, s, gy)
        for k, v in (("x", x), ("spec", spec), ("s", s), ("y", y), ("gy", gy), ("gs", gs)):
            out[f"c{i}_{k}"] = v
        metas.append([N, wl, inc, nfft, 0 if wt == "hann" else 1])
    out["c_meta"] = np.array(metas)
    # config.center = False: stft_custom passes it to torch.stft (src/evaluate.py:116); istft_custom raises for it
    metas = []
    for i, (N, n, h, w) in enumerate(((2000, 512, 128, 512), (3000, 320, 160, 320))):
        c = SimpleNamespace(n_fft=n, hop_length=h, win_length=w, center=False)
        x = torch.randn(2, 1, N, generator=g, requires_grad=True)
        spec = stft_custom(x, c)
        gspec = torch.randn(spec.shape, generator=g)
        (gx,) = torch.autograd.grad(spec, x, gspec)
        for k, v in (("x", x), ("spec", spec), ("gspec", gspec), ("gx", gx)):
            out[f"n{i}_{k}"] = v
        metas.append([N, n, h, w])
        try:
            istft_custom(spec.detach(), N, c)
            raise AssertionError("the reference was expected to raise for center=False")
        except RuntimeError:
            pass
    out["n_meta"] = np.array(metas)
    save("general_geometry", **out)


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


def _stub_missing_third_party():
    """src.distrib / src.dataset / src.utils import omegaconf, julius, librosa, ... at module level (absent here and
    unused by collate_fn_pad): any import of those names resolves to an empty stub package, for the import only."""
    import importlib.abc
    import importlib.machinery
    import types

    class Stub(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            m = Stub(self.__name__ + "." + k)
            setattr(self, k, m)
            return m

        def __call__(self, *a, **k):
            return None

    class Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        MISSING = ("omegaconf", "julius", "librosa", "pesq", "pypesq", "pystoi", "museval", "clarity", "soundfile",
                   "torchaudio", "hydra", "tensorboard")

        def find_spec(self, name, path, target=None):
            if name.split(".")[0] in self.MISSING:
                return importlib.machinery.ModuleSpec(name, self, is_package=True)

        def create_module(self, spec):
            return Stub(spec.name)

        def exec_module(self, module):
            pass

    if not any(type(f).__name__ == "Finder" for f in sys.meta_path):
        sys.meta_path.append(Finder())


def gen_losses_evaluate_collate_features():
    """The "next" rows of SURVEY 8f from the REAL reference: src.loss (si_snr, loss_sisdr, PSA: src/loss.py:14-56),
    src.evaluate.evaluate with model=None (src/evaluate.py:10-98), collate_fn_pad (src/distrib.py:38-98) and the
    models' magnitude features (Amplitude module dcunet.py:372-379; unet.py:40, dnn.py:98, crn.py:101 captured with a
    forward pre-hook on the first layer of a real model forward)."""
    from src.loss import si_snr, loss_sisdr, loss_phase_sensitive_spectral_approximation
    from src.evaluate import evaluate
    g = torch.Generator().manual_seed(4711)
    out = {}
    # --- SI-SNR: a few shapes, incl. near-silent targets
    for tag, shape in {"a": (4, 1, 3000), "b": (2, 2, 1, 1777), "c": (3, 8000)}.items():
        s2 = torch.randn(*shape, generator=g)
        s1 = s2 + 0.5 * torch.randn(*shape, generator=g)
        out[f"sisnr_s1_{tag}"], out[f"sisnr_s2_{tag}"] = s1, s2
        out[f"sisnr_{tag}"] = si_snr(s1, s2)
        out[f"sisdr_loss_{tag}"] = loss_sisdr(s1, s2)
    # --- PSA on three spectra
    enh, tgt, mix = (torch.randn(2, 1, 257, 40, 2, generator=g) for _ in range(3))
    out["psa_enh"], out["psa_tgt"], out["psa_mix"] = enh, tgt, mix
    out["psa"] = loss_phase_sensitive_spectral_approximation(enh, tgt, mix)
    save("losses", **out)

    # --- evaluate(model=None): z-score and no normalisation, lengths that do / do not need the zero-filled tail
    out = {}
    for tag, (shape, norm, n_fft, hop, segment) in {
            "zscore": ((1, 2, 5000), "z-score", 512, 128, 0.128), "plain": ((2, 1, 4096), "none", 512, 128, 0.128),
            "n1024": ((1, 1, 9000), "z-score", 1024, 256, 0.256)}.items():
        mixture = 0.3 * torch.randn(*shape, generator=g) + 0.05
        conf = SimpleNamespace(dset=SimpleNamespace(norm=norm, sample_rate=16000),
                               model=SimpleNamespace(name="unet", n_fft=n_fft, hop_length=hop, win_length=n_fft, center=True,
                                                     segment=segment, sources=["clean"]))
        out[f"mix_{tag}"] = mixture
        out[f"enh_{tag}"] = evaluate(mixture, None, "cpu", conf)
        out[f"meta_{tag}"] = np.array([n_fft, hop, int(16000 * segment), 1 if norm == "z-score" else 0])
    save("evaluate", **out)

    # --- collate_fn_pad
    _stub_missing_third_party()
    from src.distrib import collate_fn_pad
    out = {}
    conf = SimpleNamespace(segment=0.25, sample_rate=8000)          # 2000-sample segments
    lengths = [4100, 1500, 2000, 6001]
    batch = []
    for i, n in enumerate(lengths):
        mixture = torch.randn(2, n, generator=g)
        sources = torch.randn(1, 2, n, generator=g)
        batch.append((mixture, sources, {"i": i}, {"i": i}, f"clip{i}"))
        out[f"mix_{i}"], out[f"src_{i}"] = mixture, sources
    for tag, drop in (("drop", True), ("pad", False)):
        bm, bs, _, _, _, index_batch = collate_fn_pad(conf, drop_last=drop)(batch)
        out[f"batch_mix_{tag}"], out[f"batch_src_{tag}"], out[f"index_{tag}"] = bm, bs, np.array(index_batch)
    out["segment_length"] = np.array(2000)
    save("collate", **out)

    # --- magnitude features from real model forwards
    from src.model.dcunet import Amplitude
    from src.model.unet import UNet
    from src.model.dnn import DeepNeuralNetwork
    from src.model.crn import CRN
    out = {}
    spec = torch.randn(2, 1, 257, 33, 2, generator=g)
    out["spec"] = spec
    out["amplitude"] = Amplitude()(spec)

    def first_input(model, x):
        grabbed = {}

        def hook(mod, inp):
            if "x" not in grabbed:
                grabbed["x"] = inp[0].detach().clone()
        hs = [m.register_forward_pre_hook(hook) for m in model.modules() if len(list(m.children())) == 0]
        model.eval()
        try:
            with torch.no_grad():
                model(x)
        except Exception as e:          # only the first layer's input matters
            print("   (forward stopped after the first layer:", type(e).__name__, ")")
        for h in hs:
            h.remove()
        return grabbed["x"]
    out["power"] = first_input(UNet(unet_channels=1, unet_layer=4), spec).reshape(2, 1, 257, 33)
    # dnn.py:98-111: sqrt -> squeeze -> transpose(1, 2) -> reshape(batch*frame, feature) is what the first Linear sees
    out["magnitude"] = first_input(DeepNeuralNetwork(nfft=512, hidden_layer=32, dnn_ema=False), spec).reshape(2, 33, 257).transpose(1, 2).reshape(2, 1, 257, 33)
    crn_spec = torch.randn(2, 1, 161, 20, 2, generator=g)
    out["crn_spec"] = crn_spec
    out["crn"] = first_input(CRN(use_lstm=False), crn_spec)
    save("features", **out)


if __name__ == "__main__":
    if "--only-general" in sys.argv:
        gen_general_geometry()
        sys.exit(0)
    if "--only-f-rows" in sys.argv:
        gen_losses_evaluate_collate_features()
        sys.exit(0)
    gen_losses_evaluate_collate_features()
    gen_tasnet_and_metric()
    if "--only-new" in sys.argv:
        sys.exit(0)
    gen_stft_istft()
    gen_grads()
    gen_conv()
    gen_dccrn_masks()
    gen_dcunet_mask()
    gen_segments()
    gen_general_geometry()
