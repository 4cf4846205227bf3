import torch, types, sys
sys.path.insert(0, '/root/repo')
import speech_enhancement_pytorch_b200 as se
def cfg(n,h,w): return types.SimpleNamespace(n_fft=n,hop_length=h,win_length=w,center=True)
def timed(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/reps*1e3
x=torch.randn(64,1,64000).cuda()
for n,h,w in ((1024,256,1024),(1024,255,1024),(256,64,256),(512,160,400),(4096,1024,4096)):
    c=cfg(n,h,w)
    with torch.no_grad():
        spec=se.stft_custom(x,c)
        t1=timed(lambda: se.stft_custom(x,c)); t2=timed(lambda: se.istft_custom(spec,64000,c))
    print(f"n={n} hop={h} win={w}: stft {t1:.1f} us, istft {t2:.1f} us", flush=True)
st, ist = se.ConvSTFT(320,160,512,"hann","complex").cuda(), se.ConviSTFT(320,160,512,None,"hann","complex").cuda()
x16=torch.randn(16,1,64000).cuda()
with torch.no_grad():
    sp=st(x16)
    print("conv 320/160/512 16 rows: stft %.1f us istft %.1f us" % (timed(lambda: st(x16)), timed(lambda: ist(sp))))
for n, h, w in ((320, 160, 320), (400, 100, 400)):
    c = cfg(n, h, w)
    with torch.no_grad():
        spec = se.stft_custom(x, c)
        t1 = timed(lambda: se.stft_custom(x, c)); t2 = timed(lambda: se.istft_custom(spec, 64000, c))
    print(f"n={n} hop={h} win={w} (Bluestein): stft {t1:.1f} us, istft {t2:.1f} us", flush=True)
