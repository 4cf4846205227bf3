"""Pin the oracle (oracle/) against golden vectors produced by the REAL reference
(tests/golden/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import spectral_np64 as o64
from oracle import spectral_oracle as oref
from conftest import GOLDEN, golden


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    a = a.astype(np.complex128 if np.iscomplexobj(a) else np.float64)
    b = b.astype(np.complex128 if np.iscomplexobj(b) else np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)


def c2(a):
    return a[..., 0] + 1j * a[..., 1]


STFT_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "stft_*.npz")))


@pytest.mark.parametrize("name", STFT_CASES)
def test_stft_istft_oracles_match_reference(name):
    g = golden(name)
    n, h, w, length = (int(v) for v in g["meta"])
    cfg = oref.make_config(n, h, w)
    x = torch.from_numpy(g["x"])
    spec = oref.stft_custom_ref(x, cfg)
    assert spec.shape == g["spec"].shape
    assert np.array_equal(spec.numpy(), g["spec"])          # same library call -> bit exact
    y = oref.istft_custom_ref(torch.from_numpy(g["spec"]), length, cfg)
    assert np.array_equal(y.numpy(), g["y"])
    # float64 first-principles restatement
    rows = g["x"].reshape(-1, g["x"].shape[-1])
    s64 = o64.stft(rows, n, h, w)
    gs = c2(g["spec"]).reshape(rows.shape[0], n // 2 + 1, -1)
    assert rel(s64, gs) < 5e-6
    for key_s, key_y in (("spec", "y"), ("spec2", "y2")):
        if key_s not in g:
            continue
        sp = c2(g[key_s]).reshape(rows.shape[0], n // 2 + 1, -1)
        y64 = o64.istft(sp, n, h, w, length)
        assert rel(y64, g[key_y].reshape(rows.shape[0], -1)) < 5e-6


@pytest.mark.parametrize("name", ["grad_n512", "grad_n1024", "grad_n512_win400"])
def test_adjoints_match_reference_autograd(name):
    g = golden(name)
    n, h, w, N = (int(v) for v in g["meta"])
    rows = g["x"].reshape(-1, N)
    gx = o64.stft_adjoint(c2(g["gspec"]).reshape(rows.shape[0], n // 2 + 1, -1), N, n, h, w)
    assert rel(gx, g["gx"].reshape(rows.shape[0], N)) < 5e-6
    nframe = g["s"].shape[-2]
    gs = o64.istft_adjoint(g["gy"].reshape(rows.shape[0], N), nframe, n, h, w)
    assert rel(gs, c2(g["gs"]).reshape(rows.shape[0], n // 2 + 1, nframe)) < 5e-6


@pytest.mark.parametrize("name", ["conv_a", "conv_b", "conv_c"])
def test_conv_transforms_match_reference(name):
    g = golden(name)
    wl, inc, nfft, length = (int(v) for v in g["meta"])
    length = None if length < 0 else length
    x = torch.from_numpy(g["x"])
    spec = oref.conv_stft_ref(x, wl, inc, nfft, "hann")
    assert rel(spec.numpy(), g["spec"]) < 1e-6
    for ks, ky in (("spec", "y"), ("spec2", "y2")):
        y = oref.conv_istft_ref(torch.from_numpy(g[ks]), wl, inc, nfft, "hann", length)
        assert y.shape == g[ky].shape
        assert rel(y.numpy(), g[ky]) < 1e-5
    win = o64.hann_periodic(wl)
    s64 = o64.conv_stft(g["x"][:, 0], wl, inc, nfft, win)
    assert rel(s64, g["spec"]) < 5e-6
    y64 = o64.conv_istft(g["spec2"], wl, inc, nfft, win, length)
    assert rel(y64, g["y2"][:, 0]) < 2e-5


def test_conv_polar_matches_reference():
    g = golden("conv_polar")
    x = torch.from_numpy(g["x"])
    mags, phase = oref.conv_stft_ref(x, 400, 100, 512, "hann", "real")
    assert rel(mags.numpy(), g["mags"]) < 1e-6
    y = oref.conv_istft_ref(torch.from_numpy(g["mags"]), 400, 100, 512, "hann", None,
                            phase=torch.from_numpy(g["phase"]))
    assert rel(y.numpy(), g["y"]) < 1e-5


@pytest.mark.parametrize("mode", ["E", "C", "R"])
def test_mask_oracle_matches_dccrn_forward(mode):
    g = golden(f"dccrn_mask_{mode}")
    specs = torch.from_numpy(g["specs"])
    nf = specs.shape[1] // 2
    spec = torch.stack([specs[:, :nf], specs[:, nf:]], -1)
    mask = torch.stack([torch.from_numpy(g["mask_re"]), torch.from_numpy(g["mask_im"])], -1)
    out = oref.mask_apply_ref(spec, mask, mode)
    got = torch.cat([out[..., 0], out[..., 1]], 1).numpy()
    assert rel(got, g["out_spec"]) < 1e-6
    m64 = o64.mask_apply(c2(spec.numpy().astype(np.float64)), c2(mask.numpy().astype(np.float64)), mode)
    assert rel(np.concatenate([m64.real, m64.imag], 1), g["out_spec"]) < 1e-5


def test_mask_oracle_matches_dcunet_forward():
    g = golden("dcunet_mask_E")
    out = oref.mask_apply_ref(torch.from_numpy(g["spec"]), torch.from_numpy(g["raw_mask"]), "E", pre_tanh=True)
    assert rel(out.numpy(), g["out"]) < 1e-6


def test_segmenting_matches_reference():
    g = golden("segments")
    nf, stride = (int(v) for v in g["meta"])
    seg = oref.segment_ref(torch.from_numpy(g["wav"]), nf, stride)
    assert np.array_equal(seg.numpy(), g["seg"])
    back = oref.stitch_ref(seg, nf, stride, g["wav"].shape[-1])
    assert np.array_equal(back.numpy(), g["wav"])


def test_mrstft_loss_fp32_vs_f64_and_gradient():
    g = torch.Generator().manual_seed(1236)
    ref = torch.randn(2, 1, 6000, generator=g)
    est = (ref + 0.1 * torch.randn(2, 1, 6000, generator=g)).requires_grad_(True)
    loss = oref.mrstft_loss_ref(est, ref)
    (grad,) = torch.autograd.grad(loss, est)
    l64, g64 = o64.mrstft_loss(est.detach().numpy(), ref.numpy(), with_grad=True)
    assert abs(float(loss.detach()) - l64) / l64 < 1e-5
    assert rel(grad.numpy().reshape(2, -1), g64) < 1e-3   # fp32 torch itself sits ~5e-4 from f64 here


OLA_TAGS = ("half", "quarter", "coprime", "gap", "abut")


@pytest.mark.parametrize("tag", OLA_TAGS)
def test_overlap_and_add_oracle_matches_reference(tag):
    g = golden("tasnet_metric")
    out = oref.overlap_and_add_ref(torch.from_numpy(g[f"sig_{tag}"]), int(g[f"step_{tag}"]))
    assert out.shape == g[f"out_{tag}"].shape
    assert np.array_equal(out.numpy(), g[f"out_{tag}"])      # same summation order -> bit exact


def test_si_sdr_metric_oracle_matches_reference():
    g = golden("tasnet_metric")
    assert abs(oref.si_sdr_metric_ref(g["sdr_ref"], g["sdr_est"]) - float(g["sdr"])) < 1e-6
    assert abs(oref.si_sdr_metric_ref(g["sdr_ref"], 0.01 * g["sdr_est"] + 0.5) - float(g["sdr_scaled"])) < 1e-5


# ---------------------------------------------------------------- SURVEY 8f rows, pinned to the REAL reference
# (fixtures: tests/golden/make_golden.py::gen_losses_evaluate_collate_features imports src.loss, src.evaluate.evaluate,
# src.distrib.collate_fn_pad and real model forwards).  These oracle functions are restatements of torch expressions,
# so the bar is bit-exactness (same ops in the same order on the same CPU build).
def test_si_snr_and_psa_oracles_are_bit_exact_with_reference_src_loss():
    gd = golden("losses")
    for tag in "abc":
        s1, s2 = torch.from_numpy(gd[f"sisnr_s1_{tag}"]), torch.from_numpy(gd[f"sisnr_s2_{tag}"])
        got = oref.si_snr_ref(s1, s2)
        assert float(got) == float(gd[f"sisnr_{tag}"]), tag
        assert float(-got) == float(gd[f"sisdr_loss_{tag}"]), tag
    enh, tgt, mix = (torch.from_numpy(gd[k]) for k in ("psa_enh", "psa_tgt", "psa_mix"))
    assert float(oref.psa_loss_ref(enh, tgt, mix)) == float(gd["psa"])


@pytest.mark.parametrize("tag", ["zscore", "plain", "n1024"])
def test_evaluate_oracle_is_bit_exact_with_reference_evaluate(tag):
    import types
    gd = golden("evaluate")
    n_fft, hop, nfeat, zscore = (int(v) for v in gd[f"meta_{tag}"])
    conf = types.SimpleNamespace(dset=types.SimpleNamespace(norm="z-score" if zscore else "none", sample_rate=16000),
                                 model=types.SimpleNamespace(name="unet", n_fft=n_fft, hop_length=hop, win_length=n_fft, center=True,
                                                             segment=nfeat / 16000.0))
    got = oref.evaluate_ref(torch.from_numpy(gd[f"mix_{tag}"]), None, conf)
    want = gd[f"enh_{tag}"]
    assert got.shape == want.shape
    assert np.array_equal(got.numpy(), want), rel(got.numpy(), want)
    # model=None makes evaluate() a (z-score -> STFT -> iSTFT -> stitch -> de-normalise) identity up to fp32 round-off
    assert rel(want, gd[f"mix_{tag}"]) < 1e-5


@pytest.mark.parametrize("tag,drop", [("drop", True), ("pad", False)])
def test_collate_oracle_is_bit_exact_with_reference_collate_fn_pad(tag, drop):
    gd = golden("collate")
    batch = [(torch.from_numpy(gd[f"mix_{i}"]), torch.from_numpy(gd[f"src_{i}"])) for i in range(4)]
    mix, src, index = oref.collate_fn_pad_ref(batch, int(gd["segment_length"]), drop_last=drop)
    assert np.array_equal(mix.numpy(), gd[f"batch_mix_{tag}"])
    assert np.array_equal(src.numpy(), gd[f"batch_src_{tag}"])
    assert list(index) == list(gd[f"index_{tag}"])


def test_magnitude_feature_oracle_is_bit_exact_with_reference_model_forwards():
    gd = golden("features")
    spec = torch.from_numpy(gd["spec"])
    for kind in ("amplitude", "power", "magnitude"):
        assert np.array_equal(oref.magnitude_feature_ref(spec, kind).numpy(), gd[kind]), kind
    got = oref.magnitude_feature_ref(torch.from_numpy(gd["crn_spec"]), "crn").numpy()
    assert np.array_equal(np.isnan(got), np.isnan(gd["crn"]))          # sqrt(re^2 - im^2): NaN where |im| > |re|, kept
    assert np.array_equal(np.nan_to_num(got), np.nan_to_num(gd["crn"]))
