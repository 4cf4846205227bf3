"""Times the fused mask+iSTFT tail against the two-stage kernels for every mask mode (cfg2 size)."""
import json, sys, torch
sys.path.insert(0, ".")
from speech_enhancement_pytorch_b200 import _native as nv
L = nv.lib()
rows, N, n, hop = 64, 64000, 1024, 256
F, T = n // 2 + 1, 1 + N // hop
dev = torch.device("cuda")
X = torch.randn(rows, F, T, 2, device=dev); Y = torch.empty_like(X); gY = torch.empty_like(X)
y = torch.empty(rows, N, device=dev); gy = torch.randn(rows, N, device=dev)
st = torch.cuda.current_stream().cuda_stream
P = lambda t: t.data_ptr()
out = {}
def timeit(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) * 1e3 / reps, 1)
for mode, name in ((0, "real"), (1, "E"), (2, "C"), (3, "R")):
    for tanh in (0, 1):
        m = torch.randn(*((rows, F, T) if mode == 0 else (rows, F, T, 2)), device=dev); gm = torch.empty_like(m)
        cnt = rows * F * T
        def two_f():
            nv.check(L.se_mask_fwd(P(X), P(m), P(Y), cnt, mode, tanh, st))
            nv.check(L.se_istft_fwd(P(Y), P(y), rows, T, N, n, hop, n, float(n), st))
        def one_f(): nv.check(L.se_mask_istft_fwd(P(X), P(m), P(y), rows, T, N, n, hop, n, float(n), mode, tanh, st))
        def two_b():
            nv.check(L.se_istft_bwd(P(gy), P(gY), rows, T, N, n, hop, n, float(n), st))
            nv.check(L.se_mask_bwd(P(X), P(m), P(gY), P(gm), 0, cnt, mode, tanh, st))
        def one_b(): nv.check(L.se_mask_istft_bwd(P(gy), P(X), P(m), P(gm), rows, T, N, n, hop, n, float(n), mode, tanh, st))
        out[f"{name}/tanh{tanh}"] = {"fwd_two": timeit(two_f), "fwd_one": timeit(one_f), "bwd_two": timeit(two_b), "bwd_one": timeit(one_b)}
        print(name, tanh, out[f"{name}/tanh{tanh}"], flush=True)
json.dump(out, open("gpurun_out/tail_timing.json", "w"), indent=1)
