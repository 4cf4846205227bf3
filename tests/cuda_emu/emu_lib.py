"""Build + load the EMULATED C-ABI library (host pointers, fibers instead of CUDA threads).

Test infrastructure only: lets the build container (no GPU) run the real kernel sources on tiny
inputs and compare against the oracle.  The product package never imports this.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "speech_enhancement_pytorch_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libse_b200_emu.so")


def _units():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if os.path.isfile(os.path.join(CSRC, f))] + [
        os.path.join(HERE, "cuda_emu.h"), os.path.join(HERE, "cuda_emu.cpp"),
        os.path.join(ROOT, "include", "se_b200.h")]
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False):
    if not (force or _stale()):
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(BUILD, exist_ok=True)
    flags = ["-O1", "-g", "-std=c++17", "-DSE_EMULATE", "-fPIC", "-I", HERE, "-I", CSRC]

    def one(src):
        obj = os.path.join(BUILD, os.path.basename(src) + ".o")
        subprocess.run(["g++"] + flags + ["-x", "c++", "-c", src, "-o", obj], check=True)
        return obj

    with ThreadPoolExecutor(max_workers=8) as pool:
        objs = list(pool.map(one, _units() + [os.path.join(HERE, "cuda_emu.cpp")]))
    subprocess.run(["g++", "-shared", "-o", LIB] + objs, check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.se_last_error.restype = ctypes.c_char_p
        _lib.se_mrstft_workspace_bytes.restype = ctypes.c_int64
        _lib.se_mrstft_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64]
    return _lib


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def check(rc):
    if rc != 0:
        raise RuntimeError(f"se error {rc}: {lib().se_last_error().decode()}")


i64 = ctypes.c_int64
f32 = ctypes.c_float
ci = ctypes.c_int


def stft_fwd(x, n, hop, win, scale):
    rows, N = x.shape
    T = 1 + N // hop
    out = np.full((rows, n // 2 + 1, T, 2), np.nan, np.float32)
    check(lib().se_stft_fwd(ptr(x), ptr(out), i64(rows), i64(N), ci(n), ci(hop), ci(win), f32(scale), None))
    return out


def stft_nocenter_fwd(x, n, hop, win, scale):
    rows, N = x.shape
    T = 1 + (N - n) // hop
    out = np.full((rows, n // 2 + 1, T, 2), np.nan, np.float32)
    check(lib().se_stft_nocenter_fwd(ptr(x), ptr(out), i64(rows), i64(N), ci(n), ci(hop), ci(win), f32(scale), None))
    return out


def stft_nocenter_bwd(g, N, n, hop, win, scale):
    rows = g.shape[0]
    out = np.full((rows, N), np.nan, np.float32)
    check(lib().se_stft_nocenter_bwd(ptr(g), ptr(out), i64(rows), i64(N), ci(n), ci(hop), ci(win), f32(scale), ci(0), None))
    return out


def stft_bwd(g, N, n, hop, win, scale, accumulate=False, init=None):
    rows = g.shape[0]
    out = np.full((rows, N), np.nan, np.float32) if init is None else init.copy()
    check(lib().se_stft_bwd(ptr(g), ptr(out), i64(rows), i64(N), ci(n), ci(hop), ci(win), f32(scale),
                            ci(int(accumulate)), None))
    return out


def istft_fwd(spec, length, n, hop, win, scale):
    rows, F, T, _ = spec.shape
    out = np.full((rows, length), np.nan, np.float32)
    check(lib().se_istft_fwd(ptr(spec), ptr(out), i64(rows), i64(T), i64(length), ci(n), ci(hop), ci(win),
                             f32(scale), None))
    return out


def istft_bwd(gy, T, n, hop, win, scale):
    rows, length = gy.shape
    out = np.full((rows, n // 2 + 1, T, 2), np.nan, np.float32)
    check(lib().se_istft_bwd(ptr(gy), ptr(out), i64(rows), i64(T), i64(length), ci(n), ci(hop), ci(win),
                             f32(scale), None))
    return out


def mask_fwd(spec, mask, mode, pre_tanh):
    out = np.full(spec.shape, np.nan, np.float32)
    check(lib().se_mask_fwd(ptr(spec), ptr(mask), ptr(out), i64(spec.size // 2), ci(mode), ci(int(pre_tanh)), None))
    return out


def mask_bwd(spec, mask, gout, mode, pre_tanh, want_gspec=True):
    gm = np.full(mask.shape, np.nan, np.float32)
    gs = np.full(spec.shape, np.nan, np.float32) if want_gspec else None
    check(lib().se_mask_bwd(ptr(spec), ptr(mask), ptr(gout), ptr(gm), ptr(gs) if want_gspec else None,
                            i64(spec.size // 2), ci(mode), ci(int(pre_tanh)), None))
    return gm, gs


def mask_planar_fwd(specs, mre, mim, mode):
    rows, nf2, nt = specs.shape
    out = np.full(specs.shape, np.nan, np.float32)
    check(lib().se_mask_planar_fwd(ptr(specs), ptr(mre), ptr(mim), ptr(out), i64(rows), i64(nf2 // 2), i64(nt), ci(mode), None))
    return out


def mask_planar_bwd(specs, mre, mim, gout, mode, want_gspec=True):
    rows, nf2, nt = specs.shape
    gre, gim = np.full(mre.shape, np.nan, np.float32), np.full(mim.shape, np.nan, np.float32)
    gs = np.full(specs.shape, np.nan, np.float32) if want_gspec else None
    check(lib().se_mask_planar_bwd(ptr(specs), ptr(mre), ptr(mim), ptr(gout), ptr(gre), ptr(gim), ptr(gs) if want_gspec else None,
                                   i64(rows), i64(nf2 // 2), i64(nt), ci(mode), None))
    return gre, gim, gs


def mrstft_fwd(est, ref):
    rows, N = est.shape
    ws = np.zeros(lib().se_mrstft_workspace_bytes(rows, N) // 8 + 1, np.float64)
    sums = np.full(9, np.nan, np.float64)
    check(lib().se_mrstft_loss_fwd(ptr(est), ptr(ref), i64(rows), i64(N), ptr(sums), ptr(ws), None))
    loss = np.full(1, np.nan, np.float32)
    check(lib().se_mrstft_loss_value(ptr(sums), i64(rows), i64(N), ptr(loss), None))
    mrstft_fwd.workspace = ws
    return sums, float(loss[0])


def mrstft_bwd(est, ref, sums, gout=1.0):
    ws = mrstft_fwd.workspace
    rows, N = est.shape
    g = np.full((rows, N), np.nan, np.float32)
    go = np.array([gout], np.float32)
    check(lib().se_mrstft_loss_bwd(ptr(est), ptr(ws), ptr(sums), ptr(go), i64(rows), i64(rows), i64(N), ptr(g), None))
    return g


def conv_stft_fwd(x, win_len, win_inc, fft_len):
    rows, N = x.shape
    T = (N + 2 * (win_len - win_inc) - win_len) // win_inc + 1
    out = np.full((rows, 2 * (fft_len // 2 + 1), T), np.nan, np.float32)
    check(lib().se_conv_stft_fwd(ptr(x), ptr(out), i64(rows), i64(N), ci(win_len), ci(win_inc), ci(fft_len), None))
    return out


def conv_istft_fwd(spec, out_len, win_len, win_inc, fft_len):
    rows, _, T = spec.shape
    out = np.full((rows, out_len), np.nan, np.float32)
    check(lib().se_conv_istft_fwd(ptr(spec), ptr(out), i64(rows), i64(T), i64(out_len), ci(win_len), ci(win_inc),
                                  ci(fft_len), None))
    return out


def conv_istft_bwd(gy, T, win_len, win_inc, fft_len):
    rows, out_len = gy.shape
    out = np.full((rows, 2 * (fft_len // 2 + 1), T), np.nan, np.float32)
    check(lib().se_conv_istft_bwd(ptr(gy), ptr(out), i64(rows), i64(T), i64(out_len), ci(win_len), ci(win_inc),
                                  ci(fft_len), None))
    return out


def register_window(values):
    w = np.ascontiguousarray(np.asarray(values, np.float64))
    wid = lib().se_register_window(ptr(w), ci(w.shape[0]))
    assert wid > 0
    return wid


def conv_stft_fwd_w(x, win_len, win_inc, fft_len, wid):
    rows, N = x.shape
    T = (N + 2 * (win_len - win_inc) - win_len) // win_inc + 1
    out = np.full((rows, 2 * (fft_len // 2 + 1), T), np.nan, np.float32)
    check(lib().se_conv_stft_fwd_w(ptr(x), ptr(out), i64(rows), i64(N), ci(win_len), ci(win_inc), ci(fft_len), ci(wid), None))
    return out


def conv_istft_fwd_w(spec, out_len, win_len, win_inc, fft_len, wid):
    rows, _, T = spec.shape
    out = np.full((rows, out_len), np.nan, np.float32)
    check(lib().se_conv_istft_fwd_w(ptr(spec), ptr(out), i64(rows), i64(T), i64(out_len), ci(win_len), ci(win_inc),
                                    ci(fft_len), ci(wid), None))
    return out


def conv_istft_bwd_w(gy, T, win_len, win_inc, fft_len, wid):
    rows, out_len = gy.shape
    out = np.full((rows, 2 * (fft_len // 2 + 1), T), np.nan, np.float32)
    check(lib().se_conv_istft_bwd_w(ptr(gy), ptr(out), i64(rows), i64(T), i64(out_len), ci(win_len), ci(win_inc),
                                    ci(fft_len), ci(wid), None))
    return out


def polar_round_trip(spec):
    rows, nf2, T = spec.shape
    mags = np.full((rows, nf2 // 2, T), np.nan, np.float32)
    phase = np.full_like(mags, np.nan)
    check(lib().se_polar_from_planar(ptr(spec), ptr(mags), ptr(phase), i64(rows), i64(nf2 // 2), i64(T), None))
    back = np.full_like(spec, np.nan)
    check(lib().se_planar_from_polar(ptr(mags), ptr(phase), ptr(back), i64(rows), i64(nf2 // 2), i64(T), None))
    return mags, phase, back


def planar_from_polar_bwd(mags, phase, g):
    rows, nf, T = mags.shape
    gm, gp = np.full_like(mags, np.nan), np.full_like(mags, np.nan)
    check(lib().se_planar_from_polar_bwd(ptr(mags), ptr(phase), ptr(g), ptr(gm), ptr(gp), i64(rows), i64(nf), i64(T), None))
    return gm, gp


def enhance_fwd(x, mask, n, hop, win, mode, pre_tanh):
    rows, N = x.shape
    out = np.full((rows, N), np.nan, np.float32)
    check(lib().se_enhance_fwd(ptr(x), ptr(mask), ptr(out), i64(rows), i64(N), ci(n), ci(hop), ci(win), ci(mode),
                               ci(int(pre_tanh)), None))
    return out


def enhance_bwd(gy, x, mask, n, hop, win, mode, pre_tanh):
    rows, N = x.shape
    out = np.full(mask.shape, np.nan, np.float32)
    check(lib().se_enhance_bwd(ptr(gy), ptr(x), ptr(mask), ptr(out), i64(rows), i64(N), ci(n), ci(hop), ci(win),
                               ci(mode), ci(int(pre_tanh)), None))
    return out


def mask_istft_fwd(spec, mask, length, n, hop, win, scale, mode, pre_tanh):
    rows, F, T, _ = spec.shape
    out = np.full((rows, length), np.nan, np.float32)
    check(lib().se_mask_istft_fwd(ptr(spec), ptr(mask), ptr(out), i64(rows), i64(T), i64(length), ci(n), ci(hop), ci(win),
                                  f32(scale), ci(mode), ci(int(pre_tanh)), None))
    return out


def mask_istft_bwd(gy, spec, mask, n, hop, win, scale, mode, pre_tanh):
    rows, F, T, _ = spec.shape
    out = np.full(mask.shape, np.nan, np.float32)
    check(lib().se_mask_istft_bwd(ptr(gy), ptr(spec), ptr(mask), ptr(out), i64(rows), i64(T), i64(gy.shape[1]), ci(n),
                                  ci(hop), ci(win), f32(scale), ci(mode), ci(int(pre_tanh)), None))
    return out


def overlap_add_fwd(sig, step):
    rows, frames, length = sig.shape
    out = np.full((rows, step * (frames - 1) + length), np.nan, np.float32)
    check(lib().se_overlap_add_fwd(ptr(sig), ptr(out), i64(rows), i64(frames), ci(length), ci(step), None))
    return out


def overlap_add_bwd(gout, frames, length, step):
    rows = gout.shape[0]
    out = np.full((rows, frames, length), np.nan, np.float32)
    check(lib().se_overlap_add_bwd(ptr(gout), ptr(out), i64(rows), i64(frames), ci(length), ci(step), None))
    return out


def stft_segments_fwd(x, nseg, seg_stride, nsample, n, hop, win, scale):
    nclip, clip_len = x.shape
    T = 1 + nsample // hop
    out = np.full((nseg * nclip, n // 2 + 1, T, 2), np.nan, np.float32)
    check(lib().se_stft_segments_fwd(ptr(x), ptr(out), i64(nseg), i64(nclip), i64(clip_len), i64(clip_len),
                                     i64(seg_stride), i64(nsample), ci(n), ci(hop), ci(win), f32(scale), None))
    return out


def row_stats(x):
    rows, length = x.shape
    out = np.full((rows, 4), np.nan, np.float32)
    check(lib().se_row_stats(ptr(x), ptr(out), i64(rows), i64(length), i64(length), None))
    return out


def stft_segments_norm_fwd(x, stats, nseg, seg_stride, nsample, n, hop, win, scale):
    nclip, clip_len = x.shape
    T = 1 + nsample // hop
    out = np.full((nseg * nclip, n // 2 + 1, T, 2), np.nan, np.float32)
    check(lib().se_stft_segments_norm_fwd(ptr(x), ptr(out), ptr(stats) if stats is not None else None, i64(nclip), i64(nclip),
                                          i64(nseg), i64(nclip), i64(clip_len), i64(clip_len), i64(seg_stride), i64(nsample),
                                          ci(n), ci(hop), ci(win), f32(scale), None))
    return out


def stft_segments_shared_fwd(x, stats, nseg, seg_stride, nsample, n, hop, win, scale):
    nclip, clip_len = x.shape
    T = 1 + nsample // hop
    lib().se_stft_segments_scratch_bytes.restype = ctypes.c_int64
    nbytes = lib().se_stft_segments_scratch_bytes(i64(nseg), i64(nclip), i64(seg_stride), i64(nsample), ci(n), ci(hop))
    assert nbytes > 0
    scratch = np.full(nbytes // 4, np.nan, np.float32)
    out = np.full((nseg * nclip, n // 2 + 1, T, 2), np.nan, np.float32)
    check(lib().se_stft_segments_shared_fwd(ptr(x), ptr(out), ptr(stats) if stats is not None else None, i64(nclip), i64(nclip),
                                            i64(nseg), i64(nclip), i64(clip_len), i64(clip_len), i64(seg_stride), i64(nsample),
                                            ci(n), ci(hop), ci(win), f32(scale), ptr(scratch), None))
    return out


def istft_stitch_fwd(spec, stats, nseg, nclip, nfeat, stride, out_len, n, hop, win, scale, div=None, chan=None):
    T = spec.shape[-2]
    out = np.full((nclip, out_len), np.nan, np.float32)
    check(lib().se_istft_stitch_fwd(ptr(spec), ptr(out), ptr(stats) if stats is not None else None, i64(div or nclip), i64(chan or nclip),
                                    i64(nseg), i64(nclip), i64(T), i64(nfeat), i64(stride), i64(out_len), i64(out_len),
                                    ci(n), ci(hop), ci(win), f32(scale), None))
    return out


def spectral_loss(enh, target, n, hop, win, kind, gout=1.0):
    rows, N = target.shape
    lib().se_spectral_loss_workspace_bytes.restype = ctypes.c_int64
    ws = np.zeros(lib().se_spectral_loss_workspace_bytes(i64(rows), i64(N), ci(hop)) // 8 + 1, np.float64)
    total = np.full(1, np.nan, np.float64)
    check(lib().se_spectral_loss_fwd(ptr(enh), ptr(target), i64(rows), i64(N), ci(n), ci(hop), ci(win), f32(1.0 / win),
                                     ci(kind), ptr(total), ptr(ws), None))
    g = np.full(enh.shape, np.nan, np.float32)
    go = np.array([gout], np.float32)
    check(lib().se_spectral_loss_bwd(ptr(enh), ptr(target), ptr(go), i64(rows), i64(rows), i64(N), ci(n), ci(hop), ci(win),
                                     f32(1.0 / win), ci(kind), ptr(g), None))
    return float(total[0]) / (enh.size), g


def conv_mask_istft_fwd(spec, mre, mim, out_len, win_len, win_inc, fft_len, mode):
    rows, _, T = spec.shape
    out = np.full((rows, out_len), np.nan, np.float32)
    check(lib().se_conv_mask_istft_fwd(ptr(spec), ptr(mre), ptr(mim), ptr(out), i64(rows), i64(T), i64(out_len), ci(win_len),
                                       ci(win_inc), ci(fft_len), ci(mode), None))
    return out


def conv_mask_istft_bwd(gy, spec, mre, mim, win_len, win_inc, fft_len, mode):
    rows, out_len = gy.shape
    T = spec.shape[-1]
    gre, gim = np.full(mre.shape, np.nan, np.float32), np.full(mim.shape, np.nan, np.float32)
    check(lib().se_conv_mask_istft_bwd(ptr(gy), ptr(spec), ptr(mre), ptr(mim), ptr(gre), ptr(gim), i64(rows), i64(T), i64(out_len),
                                       ci(win_len), ci(win_inc), ci(fft_len), ci(mode), None))
    return gre, gim
