"""The product library builds for sm_100a, loads, and exports every symbol include/se_b200.h
declares (no compute calls: there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "se_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(se_[a-z_0-9]+)\s*\(", text)))


def test_header_and_python_binding_agree():
    from speech_enhancement_pytorch_b200 import _native
    assert sorted(_native.EXPORTS) == declared_symbols()


def test_library_builds_and_exports_all_symbols():
    from speech_enhancement_pytorch_b200 import _native
    path = _native.build()
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.se_version() >= 100


def test_sass_is_sm100():
    from speech_enhancement_pytorch_b200 import _native
    import subprocess
    path = _native.build()
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_argument_errors_do_not_need_a_gpu():
    from speech_enhancement_pytorch_b200 import _native
    L = _native.lib()
    assert L.se_stft_fwd(None, None, 1, 4096, 512, 128, 512, 1.0, None) == -1
    assert b"null" in L.se_last_error()
    buf = (ctypes.c_float * 4)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert L.se_stft_fwd(p, p, 1, 4096, 321, 80, 321, 1.0, None) == -2          # odd n_fft: clear error, no fallback
    assert L.se_stft_fwd(p, p, 1, 4096, 512, 600, 512, 1.0, None) == -2          # hop > n_fft
    assert L.se_stft_fwd(p, p, 1, 100, 512, 100, 512, 1.0, None) == -1           # general geometry: reflect padding needs N > n/2
    assert L.se_geometry_tuned(512, 128) == 1 and L.se_geometry_tuned(512, 100) == 0 and L.se_geometry_tuned(256, 64) == 0
    assert L.se_conv_geometry_tuned(400, 100, 512) == 1 and L.se_conv_geometry_tuned(320, 160, 512) == 0
    assert L.se_enhance_fwd(p, p, p, 1, 4096, 256, 64, 256, 1, 0, None) == -2     # fused ops: tuned geometries only
    assert L.se_istft_fwd(p, p, 1, 0, 100, 512, 128, 512, 1.0, None) == -1
    assert L.se_mask_fwd(p, p, p, 4, 7, 0, None) == -2


def test_torch_extension_registers_the_operators():
    """The PyTorch C++ extension (TORCH_LIBRARY(se_b200), csrc_torch/se_torch.cpp) loads without a GPU and exposes
    the step's operators; calling one on a CPU tensor raises instead of falling back."""
    import pytest
    import torch
    from speech_enhancement_pytorch_b200 import _native as nv
    ops = nv.torch_ops()
    for name in ("stft", "istft", "mask", "mask_istft", "enhance", "mrstft_loss", "conv_stft", "conv_istft", "conv_mask_istft"):
        assert hasattr(ops, name), name
    assert "Tensor x, int n_fft, int hop, int win_length, float scale" in str(torch.ops.se_b200.stft.default._schema)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.mrstft_loss(torch.zeros(1, 4096), torch.zeros(1, 4096))
    with pytest.raises(NotImplementedError):
        ops.istft(torch.zeros(1, 161, 5, 2), 640, 321, 80, 321, 321.0)
