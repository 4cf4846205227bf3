import torch, types, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_enhancement_pytorch_b200 as se
c = types.SimpleNamespace(n_fft=1024, hop_length=255, win_length=1024, center=True)
x = torch.randn(64, 1, 64000).cuda()
with torch.no_grad():
    for _ in range(3):
        spec = se.stft_custom(x, c)
        y = se.istft_custom(spec, 64000, c)
torch.cuda.synchronize()
