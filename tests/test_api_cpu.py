"""Host-side logic of the reference-facing API that can be checked without a GPU."""
import types

import pytest
import torch

import speech_enhancement_pytorch_b200 as se


def cfg(n=512, h=128, w=512, center=True):
    return types.SimpleNamespace(n_fft=n, hop_length=h, win_length=w, center=center)


def test_cpu_tensors_are_refused_loudly():
    x = torch.randn(1, 1, 4096)
    with pytest.raises(RuntimeError, match="CUDA"):
        se.stft_custom(x, cfg())
    with pytest.raises(RuntimeError, match="CUDA"):
        se.istft_custom(torch.randn(1, 1, 257, 33, 2), 4096, cfg())
    with pytest.raises(RuntimeError, match="CUDA"):
        se.loss_mrstft(x, x.clone())
    with pytest.raises(RuntimeError, match="CUDA"):
        se.apply_mask(torch.randn(1, 1, 257, 33, 2), torch.randn(1, 1, 257, 33, 2), "E")


def test_unsupported_configs_raise_not_fallback():
    x = torch.randn(1, 1, 4096)
    with pytest.raises(NotImplementedError):
        se.stft_custom(x, cfg(321, 80, 321))          # odd n_fft
    with pytest.raises(RuntimeError, match="CUDA"):
        se.stft_custom(x, cfg(320, 160, 320))         # the commented CRN setting (src/conf/config.yaml:78-80) is built, CUDA-only
    with pytest.raises(NotImplementedError):
        se.stft_custom(x, cfg(512, 600, 512))         # hop > n_fft
    with pytest.raises(RuntimeError, match="CUDA"):
        se.stft_custom(x, cfg(512, 100, 512))         # a general geometry is built -- but there is still no CPU path
    with pytest.raises(RuntimeError, match="CUDA"):
        se.stft_custom(x, cfg(center=False))          # built (no padding), CUDA-only
    with pytest.raises(RuntimeError, match="overlap add"):
        se.istft_custom(torch.randn(1, 1, 257, 33, 2), 4096, cfg(center=False))     # the reference's torch.istft raises too
    with pytest.raises(ValueError):
        se.stft_custom(torch.randn(4096), cfg())
    with pytest.raises(ValueError):
        se.istft_custom(torch.randn(257, 33, 2), 4096, cfg())
    with pytest.raises(ValueError):
        se.apply_mask(torch.randn(1, 1, 257, 33, 2), torch.randn(1, 1, 257, 33), "E")
    with pytest.raises(ValueError):
        se.apply_mask(torch.randn(1, 1, 257, 33, 2), torch.randn(1, 1, 257, 33, 2), "Z")


def test_dccrn_modules_load_a_reference_state_dict_strictly():
    """Same registered buffers as the reference's ConvSTFT / ConviSTFT (src/model/dccrn.py:681,714,720-721), built like
    `init_kernels` (:649-666) builds them, for every window type scipy knows -- including the constructors' default
    'hamming': a reference checkpoint loads with strict=True.  The fixture values are the REAL reference's buffers
    (tests/golden/conv_*.npz hold hann; the hamming ones are rebuilt here with the reference's own formula)."""
    import numpy as np
    from scipy.signal import get_window
    st, ist = se.ConvSTFT(400, 100, 512), se.ConviSTFT(400, 100, 512)            # defaults: win_type='hamming'
    assert set(st.state_dict()) == {"weight"} and set(ist.state_dict()) == {"weight", "window", "enframe"}
    assert tuple(st.weight.shape) == (514, 1, 400) and tuple(ist.weight.shape) == (514, 1, 400)
    assert tuple(ist.window.shape) == (1, 400, 1) and tuple(ist.enframe.shape) == (400, 1, 400)
    w = get_window("hamming", 400, fftbins=True)
    basis = np.fft.rfft(np.eye(512))[:400]
    kernel = np.concatenate([np.real(basis), np.imag(basis)], 1).T
    assert np.array_equal(st.weight.numpy(), (kernel * w)[:, None, :].astype(np.float32))
    assert np.array_equal(ist.weight.numpy(), (np.linalg.pinv(kernel).T * w)[:, None, :].astype(np.float32))
    fresh = se.ConviSTFT(400, 100, 512, win_type="hann")
    fresh.load_state_dict(ist.state_dict(), strict=True)
    for wt in ("hann", "hamming", "blackman", None, "None"):
        se.ConvSTFT(400, 100, 512, wt)


def test_modules_keep_reference_attributes():
    st = se.ConvSTFT(400, 100, None, "hann", "complex")
    assert (st.fft_len, st.stride, st.win_len, st.dim) == (512, 100, 400, 512)
    ist = se.ConviSTFT(400, 100, 512, 16384, "hann", "complex")
    assert (ist.length, ist.stride, ist.win_len) == (16384, 100, 400)


def test_fused_tail_and_tasnet_entry_points_validate_without_a_gpu():
    spec, mask = torch.randn(1, 1, 257, 33, 2), torch.randn(1, 1, 257, 33, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        se.apply_mask_istft(spec, mask, 4096, cfg(), "E", True)
    with pytest.raises(ValueError):
        se.apply_mask_istft(spec, mask[..., 0], 4096, cfg(), "E")            # complex modes need [...,F,T,2]
    with pytest.raises(ValueError):
        se.apply_mask_istft(spec, mask, 4096, cfg(), "Z")
    with pytest.raises(ValueError):
        se.apply_mask_istft(spec[0, 0], mask[0, 0], 4096, cfg(), "C")        # 5-D / 6-D only, like istft_custom
    with pytest.raises(RuntimeError):
        se.apply_mask_istft(torch.randn(1, 1, 129, 33, 2), torch.randn(1, 1, 129, 33, 2), 4096, cfg(), "C")
    with pytest.raises(NotImplementedError):
        se.apply_mask_istft(spec, mask, 4096, cfg(512, 600, 512), "C")         # hop > n_fft
    with pytest.raises(RuntimeError, match="CUDA"):
        se.apply_mask_istft(spec, mask, 4096, cfg(512, 100, 512), "C")        # general geometry: two stages, still CUDA-only
    with pytest.raises(RuntimeError, match="CUDA"):
        se.overlap_and_add(torch.randn(2, 5, 40), 20)
    with pytest.raises(ValueError):
        se.overlap_and_add(torch.randn(40), 20)
    with pytest.raises(ValueError):
        se.overlap_and_add(torch.randn(2, 5, 40), 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        se.SI_SDR(torch.randn(2, 100), torch.randn(2, 100))
    with pytest.raises(ValueError):
        se.SI_SDR(torch.randn(2, 100), torch.randn(2, 99))


def test_dccrn_tail_entry_points_validate_without_a_gpu():
    specs, mre, mim = torch.randn(2, 514, 9), torch.randn(2, 257, 9), torch.randn(2, 257, 9)
    with pytest.raises(RuntimeError, match="CUDA"):
        se.apply_mask_dccrn(specs, mre, mim, "E")
    with pytest.raises(ValueError):
        se.apply_mask_dccrn(specs, mre, mim, "real")                       # DCCRN has E / C / R only (dccrn.py:203-221)
    with pytest.raises(ValueError):
        se.apply_mask_dccrn(specs, mre[:, 1:], mim[:, 1:], "E")            # masks must be padded at DC (dccrn.py:200-201)
    with pytest.raises(ValueError):
        se.apply_mask_dccrn(torch.randn(2, 513, 9), mre, mim, "C")
    ist = se.ConviSTFT(400, 100, 512, 600, "hann", "complex")
    with pytest.raises(RuntimeError, match="CUDA"):
        ist.forward_masked(specs, mre, mim, "E")
    with pytest.raises(ValueError):
        ist.forward_masked(specs, mre, mim, "Z")
    with pytest.raises(ValueError):
        ist.forward_masked(specs, mre[:, 1:], mim, "E")
    assert ist._out_len(9) == 600 and se.ConviSTFT(400, 100, 512, None, "hann", "complex")._out_len(9) == 600


def test_peer_exchange_binding_validates_without_a_gpu():
    import ctypes
    from speech_enhancement_pytorch_b200 import _native as nv
    L = nv.lib()
    buf = (ctypes.c_double * 9)()
    ptrs = (ctypes.c_void_p * 2)(ctypes.addressof(buf), None)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert L.se_mrstft_exchange_value(None, ptrs, 2, 0, 4, 4096, None, None) == -1
    assert L.se_mrstft_exchange_value(p, ptrs, 17, 0, 4, 4096, None, None) == -1       # at most 16 ranks
    assert L.se_mrstft_exchange_value(p, ptrs, 2, 2, 4, 4096, None, None) == -1        # rank out of range
    assert L.se_mrstft_exchange_value(p, ptrs, 2, 0, 4, 4096, None, None) == -1        # a peer buffer is missing
    assert b"exchange buffer" in L.se_last_error()
    assert L.se_mask_planar_fwd(p, p, p, p, 1, 4, 4, 0, None) == -2                     # REAL is not a DCCRN mode
    assert L.se_conv_mask_istft_fwd(p, p, p, p, 1, 9, 600, 400, 100, 512, 7, None) == -2
