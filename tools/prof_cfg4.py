import sys, torch
sys.path.insert(0, '/root/repo')
import speech_enhancement_pytorch_b200 as se
dev='cuda'; N=64000
st, ist = se.ConvSTFT(400, 100, 512, "hann", "complex"), se.ConviSTFT(400, 100, 512, N, "hann", "complex")
x=torch.randn(16,1,N,device=dev); mre=torch.randn(16,257,643,device=dev,requires_grad=True); mim=torch.randn(16,257,643,device=dev,requires_grad=True)
for _ in range(3):
    mre.grad=None; mim.grad=None
    y = ist.forward_masked(st(x), mre, mim, "E"); y.backward(y)
torch.cuda.synchronize()
