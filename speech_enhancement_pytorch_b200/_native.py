"""ctypes binding of libse_b200.so (the C-ABI in include/se_b200.h).

There is NO fallback: if the library is missing or the tensors are not CUDA fp32, the calls
raise.  PyTorch is used only for device memory, streams and autograd plumbing.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libse_b200.so")
TORCH_LIB_PATH = os.path.join(_PKG, "libse_b200_torch.so")      # TORCH_LIBRARY(se_b200) + C++ autograd over the C-ABI
CSRC = os.path.join(_PKG, "csrc")
CSRC_TORCH = os.path.join(_PKG, "csrc_torch")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

EXPORTS = (
    "se_version", "se_last_error", "se_geometry_tuned", "se_conv_geometry_tuned", "se_stft_nocenter_fwd", "se_stft_nocenter_bwd", "se_stft_fwd", "se_stft_segments_fwd", "se_row_stats", "se_stft_segments_norm_fwd", "se_istft_stitch_fwd",
    "se_stft_segments_scratch_bytes", "se_stft_segments_shared_fwd", "se_magnitude_feature", "se_stft_feature_fwd", "se_stft_bwd", "se_istft_fwd", "se_istft_bwd",
    "se_mask_fwd", "se_mask_bwd", "se_mrstft_workspace_bytes", "se_mrstft_loss_fwd",
    "se_mrstft_loss_value", "se_mrstft_loss_bwd", "se_spectral_loss_workspace_bytes", "se_spectral_loss_fwd",
    "se_spectral_loss_bwd", "se_sisnr_fwd", "se_sisnr_bwd", "se_psa_workspace_bytes", "se_psa_loss_fwd", "se_psa_loss_bwd", "se_enhance_fwd", "se_enhance_bwd",
    "se_mask_istft_fwd", "se_mask_istft_bwd", "se_overlap_add_fwd", "se_overlap_add_bwd",
    "se_conv_stft_fwd", "se_conv_istft_fwd", "se_conv_istft_bwd",
    "se_mask_planar_fwd", "se_mask_planar_bwd", "se_conv_mask_istft_fwd", "se_conv_mask_istft_bwd",
    "se_p2p_create", "se_p2p_open", "se_p2p_close", "se_p2p_destroy", "se_mrstft_exchange_value",
    "se_mrstft_exchange_rows_value", "se_mrstft_loss_value_dev", "se_mrstft_loss_fwd_value",
    "se_register_window", "se_conv_stft_fwd_w", "se_conv_istft_fwd_w", "se_conv_istft_bwd_w", "se_conv_mask_istft_fwd_w",
    "se_conv_mask_istft_bwd_w", "se_polar_from_planar", "se_planar_from_polar", "se_planar_from_polar_bwd",
)

MASK_MODES = {"real": 0, "E": 1, "C": 2, "R": 3}
FEATURE_KINDS = {"power": 0, "magnitude": 1, "amplitude": 2, "crn": 3}


BUILD_DIR = os.path.join(_PKG, "build")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if os.path.isfile(os.path.join(CSRC, f)))


def _units():
    return [p for p in _sources() if p.endswith(".cu")]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _sources() + [os.path.join(_ROOT, "include", "se_b200.h")]
    return any(os.path.getmtime(p) > t for p in deps if os.path.exists(p))


ALT_LIB_PATH = os.path.join(_PKG, "libse_alt_tc.so")      # measured alternative (tensor-core DFT-as-GEMM); never loaded by the package
CSRC_ALT = os.path.join(_PKG, "csrc_alt")


def build_alt(force: bool = False) -> str:
    """Compile csrc_alt/*.cu (the measured tensor-core alternative of profiles/r02_notes.md (c), used only by
    tools/tc_dft_bench.py) for sm_100a so that every CUDA source in the tree is covered by the build check."""
    srcs = sorted(os.path.join(CSRC_ALT, f) for f in os.listdir(CSRC_ALT) if f.endswith(".cu"))
    if not force and os.path.exists(ALT_LIB_PATH) and all(os.path.getmtime(p) <= os.path.getmtime(ALT_LIB_PATH) for p in srcs):
        return ALT_LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    res = subprocess.run([nvcc] + NVCC_FLAGS + ["-o", ALT_LIB_PATH] + srcs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc (csrc_alt) failed:\n" + res.stdout + res.stderr)
    return ALT_LIB_PATH


def _torch_sources():
    return sorted(os.path.join(CSRC_TORCH, f) for f in os.listdir(CSRC_TORCH) if f.endswith(".cpp"))


def needs_torch_build() -> bool:
    if not os.path.exists(TORCH_LIB_PATH):
        return True
    t = os.path.getmtime(TORCH_LIB_PATH)
    deps = _torch_sources() + [os.path.join(_ROOT, "include", "se_b200.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build_torch_ext(force: bool = False) -> str:
    """g++ the PyTorch C++ extension (csrc_torch/*.cpp: operator registrations + C++ autograd nodes, no kernels)
    against the torch headers and link it to the in-tree libse_b200.so."""
    if not (force or needs_torch_build()):
        return TORCH_LIB_PATH
    from torch.utils import cpp_extension as ce
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared",
           "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI))]
    cmd += [f"-I{p}" for p in ce.include_paths()] + [f"-I{cuda_inc}"] + _torch_sources()
    cmd += ["-o", TORCH_LIB_PATH, f"-L{_PKG}", "-lse_b200", f"-L{libdir}", "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda",
            "-ltorch_cuda", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ (torch extension) failed:\n" + res.stdout + res.stderr)
    return TORCH_LIB_PATH


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a with nvcc (one process per translation unit, in parallel) and
    link the in-tree libse_b200.so."""
    if not (force or needs_build()):
        build_torch_ext(force=False)
        build_alt(force=False)
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(BUILD_DIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".o")
        res = subprocess.run([nvcc] + flags + ["-c", "-o", obj, src], capture_output=True, text=True)
        return obj, res

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, _units()))
    log = ""
    for obj, res in results:
        log += res.stderr
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    res = subprocess.run([nvcc, "-shared", "-o", LIB_PATH] + [o for o, _ in results], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    if verbose:
        with open(os.path.join(BUILD_DIR, "ptxas.log"), "w") as f:
            f.write(log)
    build_torch_ext(force=True)
    build_alt(force=force)
    return LIB_PATH


_lib = None
_lock = threading.Lock()
_c = ctypes
_I64, _INT, _F32, _PTR = _c.c_int64, _c.c_int, _c.c_float, _c.c_void_p


def lib():
    """Load (once) and return the C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback.")
            L = ctypes.CDLL(LIB_PATH)
            L.se_last_error.restype = _c.c_char_p
            L.se_mrstft_workspace_bytes.restype = _I64
            L.se_mrstft_workspace_bytes.argtypes = [_I64, _I64]
            L.se_stft_fwd.argtypes = [_PTR, _PTR, _I64, _I64, _INT, _INT, _INT, _F32, _PTR]
            L.se_stft_segments_fwd.argtypes = [_PTR, _PTR, _I64, _I64, _I64, _I64, _I64, _I64, _INT, _INT, _INT, _F32, _PTR]
            L.se_geometry_tuned.argtypes = [_INT, _INT]
            L.se_conv_geometry_tuned.argtypes = [_INT, _INT, _INT]
            L.se_row_stats.argtypes = [_PTR, _PTR, _I64, _I64, _I64, _PTR]
            L.se_stft_segments_norm_fwd.argtypes = [_PTR, _PTR, _PTR, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _INT, _INT, _INT, _F32, _PTR]
            L.se_stft_segments_scratch_bytes.restype = _I64
            L.se_stft_segments_scratch_bytes.argtypes = [_I64, _I64, _I64, _I64, _INT, _INT]
            L.se_stft_segments_shared_fwd.argtypes = [_PTR, _PTR, _PTR, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _INT, _INT, _INT, _F32, _PTR, _PTR]
            L.se_istft_stitch_fwd.argtypes = [_PTR, _PTR, _PTR, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _INT, _INT, _INT, _F32, _PTR]
            L.se_magnitude_feature.argtypes = [_PTR, _PTR, _I64, _INT, _PTR]
            L.se_stft_feature_fwd.argtypes = [_PTR, _PTR, _PTR, _I64, _I64, _INT, _INT, _INT, _F32, _INT, _PTR]
            L.se_stft_bwd.argtypes = [_PTR, _PTR, _I64, _I64, _INT, _INT, _INT, _F32, _INT, _PTR]
            L.se_istft_fwd.argtypes = [_PTR, _PTR, _I64, _I64, _I64, _INT, _INT, _INT, _F32, _PTR]
            L.se_istft_bwd.argtypes = [_PTR, _PTR, _I64, _I64, _I64, _INT, _INT, _INT, _F32, _PTR]
            L.se_mask_fwd.argtypes = [_PTR, _PTR, _PTR, _I64, _INT, _INT, _PTR]
            L.se_mask_bwd.argtypes = [_PTR, _PTR, _PTR, _PTR, _PTR, _I64, _INT, _INT, _PTR]
            L.se_mrstft_loss_fwd.argtypes = [_PTR, _PTR, _I64, _I64, _PTR, _PTR, _PTR]
            L.se_mrstft_loss_value.argtypes = [_PTR, _I64, _I64, _PTR, _PTR]
            L.se_mrstft_loss_fwd_value.argtypes = [_PTR, _PTR, _I64, _I64, _PTR, _PTR, _PTR, _PTR]
            L.se_mrstft_loss_bwd.argtypes = [_PTR, _PTR, _PTR, _PTR, _I64, _I64, _I64, _PTR, _PTR]
            L.se_spectral_loss_workspace_bytes.restype = _I64
            L.se_spectral_loss_workspace_bytes.argtypes = [_I64, _I64, _INT]
            L.se_spectral_loss_fwd.argtypes = [_PTR, _PTR, _I64, _I64, _INT, _INT, _INT, _F32, _INT, _PTR, _PTR, _PTR]
            L.se_spectral_loss_bwd.argtypes = [_PTR, _PTR, _PTR, _I64, _I64, _I64, _INT, _INT, _INT, _F32, _INT, _PTR, _PTR]
            L.se_sisnr_fwd.argtypes = [_PTR, _PTR, _I64, _I64, _c.c_double, _PTR, _PTR, _PTR]
            L.se_sisnr_bwd.argtypes = [_PTR, _PTR, _PTR, _PTR, _F32, _I64, _I64, _c.c_double, _PTR, _PTR]
            L.se_psa_workspace_bytes.restype = _I64
            L.se_psa_workspace_bytes.argtypes = [_I64]
            L.se_psa_loss_fwd.argtypes = [_PTR, _PTR, _PTR, _I64, _PTR, _PTR, _PTR]
            L.se_psa_loss_bwd.argtypes = [_PTR, _PTR, _PTR, _PTR, _I64, _I64, _PTR, _PTR]
            L.se_enhance_fwd.argtypes = [_PTR, _PTR, _PTR, _I64, _I64, _INT, _INT, _INT, _INT, _INT, _PTR]
            L.se_enhance_bwd.argtypes = [_PTR, _PTR, _PTR, _PTR, _I64, _I64, _INT, _INT, _INT, _INT, _INT, _PTR]
            L.se_mask_istft_fwd.argtypes = [_PTR, _PTR, _PTR, _I64, _I64, _I64, _INT, _INT, _INT, _F32, _INT, _INT, _PTR]
            L.se_mask_istft_bwd.argtypes = [_PTR, _PTR, _PTR, _PTR, _I64, _I64, _I64, _INT, _INT, _INT, _F32, _INT, _INT, _PTR]
            L.se_overlap_add_fwd.argtypes = [_PTR, _PTR, _I64, _I64, _INT, _INT, _PTR]
            L.se_overlap_add_bwd.argtypes = [_PTR, _PTR, _I64, _I64, _INT, _INT, _PTR]
            L.se_conv_stft_fwd.argtypes = [_PTR, _PTR, _I64, _I64, _INT, _INT, _INT, _PTR]
            L.se_conv_istft_fwd.argtypes = [_PTR, _PTR, _I64, _I64, _I64, _INT, _INT, _INT, _PTR]
            L.se_conv_istft_bwd.argtypes = [_PTR, _PTR, _I64, _I64, _I64, _INT, _INT, _INT, _PTR]
            L.se_register_window.argtypes = [_PTR, _INT]
            L.se_conv_stft_fwd_w.argtypes = [_PTR, _PTR, _I64, _I64, _INT, _INT, _INT, _INT, _PTR]
            L.se_conv_istft_fwd_w.argtypes = [_PTR, _PTR, _I64, _I64, _I64, _INT, _INT, _INT, _INT, _PTR]
            L.se_conv_istft_bwd_w.argtypes = [_PTR, _PTR, _I64, _I64, _I64, _INT, _INT, _INT, _INT, _PTR]
            L.se_conv_mask_istft_fwd_w.argtypes = [_PTR, _PTR, _PTR, _PTR, _I64, _I64, _I64, _INT, _INT, _INT, _INT, _INT, _PTR]
            L.se_conv_mask_istft_bwd_w.argtypes = [_PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _I64, _I64, _I64, _INT, _INT, _INT, _INT, _INT, _PTR]
            L.se_polar_from_planar.argtypes = [_PTR, _PTR, _PTR, _I64, _I64, _I64, _PTR]
            L.se_planar_from_polar.argtypes = [_PTR, _PTR, _PTR, _I64, _I64, _I64, _PTR]
            L.se_planar_from_polar_bwd.argtypes = [_PTR, _PTR, _PTR, _PTR, _PTR, _I64, _I64, _I64, _PTR]
            L.se_mask_planar_fwd.argtypes = [_PTR, _PTR, _PTR, _PTR, _I64, _I64, _I64, _INT, _PTR]
            L.se_mask_planar_bwd.argtypes = [_PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _I64, _I64, _I64, _INT, _PTR]
            L.se_conv_mask_istft_fwd.argtypes = [_PTR, _PTR, _PTR, _PTR, _I64, _I64, _I64, _INT, _INT, _INT, _INT, _PTR]
            L.se_conv_mask_istft_bwd.argtypes = [_PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _I64, _I64, _I64, _INT, _INT, _INT, _INT, _PTR]
            L.se_p2p_create.argtypes = [_c.POINTER(_PTR), _c.c_char_p]
            L.se_p2p_open.argtypes = [_c.c_char_p, _c.POINTER(_PTR)]
            L.se_p2p_close.argtypes = [_PTR]
            L.se_p2p_destroy.argtypes = [_PTR]
            L.se_mrstft_exchange_value.argtypes = [_PTR, _c.POINTER(_PTR), _INT, _INT, _I64, _I64, _PTR, _PTR]
            L.se_mrstft_exchange_rows_value.argtypes = [_PTR, _c.POINTER(_PTR), _INT, _INT, _I64, _PTR, _PTR]
            L.se_mrstft_loss_value_dev.argtypes = [_PTR, _I64, _PTR, _PTR]
            _lib = L
    return _lib


_torch_ops = None


def torch_ops():
    """torch.ops.se_b200 (the C++ extension).  Fails loudly when it has not been built: there is no Python or
    PyTorch fallback for these operators."""
    global _torch_ops
    if _torch_ops is None:
        lib()                                           # libse_b200.so first (the extension links against it)
        with _lock:
            if _torch_ops is None:
                if not os.path.exists(TORCH_LIB_PATH):
                    raise RuntimeError(
                        f"{TORCH_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`.  "
                        "This package has no CPU or PyTorch fallback.")
                torch.ops.load_library(TORCH_LIB_PATH)
                _torch_ops = torch.ops.se_b200
    return _torch_ops


class SEError(RuntimeError):
    pass


_ERR = {-1: ValueError, -2: NotImplementedError, -3: SEError, -4: RuntimeError}


def check(rc: int):
    if rc != 0:
        msg = lib().se_last_error().decode()
        raise _ERR.get(rc, SEError)(f"se_b200[{rc}]: {msg}")


def require_cuda_f32(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("speech_enhancement_pytorch_b200 runs on CUDA tensors only "
                               "(hand-written sm_100a kernels; there is no CPU fallback)")
        if t.dtype != torch.float32:
            raise TypeError(f"expected float32 tensors, got {t.dtype}")
        if not t.is_contiguous():
            raise ValueError("internal error: tensor must be contiguous at the C-ABI")


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class on_device:
    """Make `device` current around a C-ABI call (tables are cached per current device).  A no-op when it
    already is -- the common single-GPU-per-process case -- because the guard costs several microseconds."""

    __slots__ = ("guard",)

    def __init__(self, device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        self.guard = None if idx == torch.cuda.current_device() else torch.cuda.device(idx)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *a):
        if self.guard is not None:
            return self.guard.__exit__(*a)
        return False
