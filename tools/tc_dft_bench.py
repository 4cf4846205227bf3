"""DFT-as-GEMM on the tensor pipe (csrc_alt/tc_dft512.cu, mma.sync TF32, 3xTF32 split) against the SIMT FFT kernel
(se_stft_fwd) for n_fft = 512 / hop 128: accuracy vs a float64 FFT and CUDA-event times.  `ncu` wraps this script for
sm__pipe_tensor_cycles_active (tools/run_gpu_tc.sh)."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speech_enhancement_pytorch_b200 import _native as nv  # noqa: E402

ALT = ctypes.CDLL(os.path.join(ROOT, "speech_enhancement_pytorch_b200", "libse_alt_tc.so"))
ALT.se_alt_tc_stft512.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_float,
                                  ctypes.c_int, ctypes.c_void_p]


def timed(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def f64_stft(x, n, hop):
    p = np.pad(x.astype(np.float64), [(0, 0), (n // 2, n // 2)], mode="reflect")
    T = 1 + x.shape[1] // hop
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / n)
    frames = np.stack([p[:, t * hop:t * hop + n] for t in range(T)], 1) * w
    return np.fft.rfft(frames, axis=-1).transpose(0, 2, 1) / n            # [rows, F, T]


def main():
    n, hop, N = 512, 128, 64000
    L = nv.lib()
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for rows in (16, 64):
        x = torch.randn(rows, N, device="cuda")
        T = 1 + N // hop
        X_simt = torch.empty(rows, n // 2 + 1, T, 2, device="cuda")
        X_tc = torch.empty_like(X_simt)
        X_tc1 = torch.empty_like(X_simt)
        simt = lambda: nv.check(L.se_stft_fwd(x.data_ptr(), X_simt.data_ptr(), rows, N, n, hop, n, 1.0 / n, st))
        tc3 = lambda: ALT.se_alt_tc_stft512(x.data_ptr(), X_tc.data_ptr(), rows, N, n, 1.0 / n, 3, st)
        tc1 = lambda: ALT.se_alt_tc_stft512(x.data_ptr(), X_tc1.data_ptr(), rows, N, n, 1.0 / n, 1, st)
        simt(); assert tc3() == 0 and tc1() == 0
        torch.cuda.synchronize()
        want = f64_stft(x[:2].cpu().numpy(), n, hop)
        def err(X):
            got = X[:2].cpu().numpy().astype(np.float64)
            got = got[..., 0] + 1j * got[..., 1]
            return float(np.abs(got - want).max() / np.abs(want).max())
        flops = rows * T * 2.0 * n * n
        r = {"simt_us": timed(simt), "tc_3xtf32_us": timed(tc3), "tc_1xtf32_us": timed(tc1),
             "simt_max_rel_err": err(X_simt), "tc_3xtf32_max_rel_err": err(X_tc), "tc_1xtf32_max_rel_err": err(X_tc1)}
        r["tc_3xtf32_tensor_tflops"] = 3 * flops / r["tc_3xtf32_us"] * 1e-6
        r["tc_1xtf32_tensor_tflops"] = flops / r["tc_1xtf32_us"] * 1e-6
        out[f"rows{rows}"] = r
    print(json.dumps(out))


if __name__ == "__main__":
    main()
