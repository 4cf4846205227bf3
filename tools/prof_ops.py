"""One launch of each standalone op at BASELINE cfg2 shapes (64 x 4 s, n_fft 1024 / hop 256) plus the MR-STFT loss,
straight through the C-ABI -- the command ncu wraps (tools/run_gpu_prof2.sh).  SE_ENGINE=1 selects the scalar engine."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_enhancement_pytorch_b200 import _native as nv  # noqa: E402

rows, N, n, hop = int(os.environ.get("ROWS", "64")), 64000, int(os.environ.get("NFFT", "1024")), 0
hop = n // 4
reps = int(os.environ.get("REPS", "2"))
F, T = n // 2 + 1, 1 + N // hop
dev = torch.device("cuda", 0)
L = nv.lib()
g = torch.Generator().manual_seed(0)
x = torch.randn(rows, N, generator=g).to(dev)
clean = (x.cpu() + 0.3 * torch.randn(rows, N, generator=g)).to(dev)
X = torch.empty(rows, F, T, 2, device=dev)
y = torch.empty(rows, N, device=dev)
gy = torch.empty(rows, N, device=dev)
gX = torch.empty_like(X)
ws = torch.empty(max(int(L.se_mrstft_workspace_bytes(rows, N)), 8), dtype=torch.uint8, device=dev)
sums = torch.empty(9, dtype=torch.float64, device=dev)
loss = torch.empty((), device=dev)
one = torch.ones((), device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
P = lambda t: t.data_ptr()
for _ in range(reps):
    nv.check(L.se_stft_fwd(P(x), P(X), rows, N, n, hop, n, 1.0 / n, st))
    nv.check(L.se_istft_fwd(P(X), P(y), rows, T, N, n, hop, n, float(n), st))
    nv.check(L.se_istft_bwd(P(y), P(gX), rows, T, N, n, hop, n, float(n), st))
    nv.check(L.se_stft_bwd(P(gX), P(gy), rows, N, n, hop, n, 1.0 / n, 0, st))
    nv.check(L.se_mrstft_loss_fwd(P(y), P(clean), rows, N, P(sums), P(ws), st))
    nv.check(L.se_mrstft_loss_value(P(sums), rows, N, P(loss), st))
    nv.check(L.se_mrstft_loss_bwd(P(y), P(ws), P(sums), P(one), rows, rows, N, P(gy), st))
torch.cuda.synchronize()
print("loss", float(loss))
