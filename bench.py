#!/usr/bin/env python
"""bench.py -- headline benchmark of the spectral hot path (BASELINE.json metric).

One "step" = one pass of the whole path over one synthetic batch, BASELINE config[1]:
  mixture [64,1,64000] (16 kHz, 4 s) -> STFT (n_fft 1024, hop 256, Hann) -> DCUnet complex-ratio
  mask (polar 'E', tanh-squashed raw mask) -> iSTFT -> multi-resolution STFT loss (512/1024/2048)
  vs the clean target -> backward to the raw mask.
metric = audio-seconds processed per second (clips * N / 16000 per step).

  python bench.py [--gpus N] [--steps K] [--warmup W]        ours (one process per GPU)
  python bench.py --impl reference ...                        the reference's CPU path (oracle port)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 16000
N_FFT, HOP, WIN = 1024, 256, 1024
RES = ((512, 128), (1024, 256), (2048, 512))
METRIC = "audio-sec/s STFT+mask+iSTFT+MR-STFT loss fwd+bwd"
UNIT = "audio-s/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=64, help="utterances per GPU per step")
    ap.add_argument("--nsample", type=int, default=64000)
    ap.add_argument("--cpu-rows", type=int, default=16, help="rows of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true", help="skip the per-kernel timing loops (for ncu runs)")
    ap.add_argument("--rounds", type=int, default=5, help="timed rounds of --steps steps; the median round is reported")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (cfg1/3/4/5)")
    ap.add_argument("--no-incumbent", action="store_true", help="skip the torch + cuFFT incumbent of the same step")
    ap.add_argument("--fused", type=int, default=-1,
                    help="composition: 2 (default, -1): stft_custom + apply_mask_istft (tail fused with the iSTFT); "
                         "0: five drop-in kernels; 1: se.enhance (everything fused, STFT recomputed in backward)")
    return ap.parse_args()


def workload_config(args, world, exchange="none (1 GPU)"):
    return {"workload": f"cfg2: DCUnet 'E' complex-ratio mask, n_fft={N_FFT} hop={HOP} Hann, "
                        f"{args.rows}x{args.nsample / SR:g} s @16 kHz per GPU, MR-STFT loss 512/1024/2048, fwd+bwd to the raw mask",
            "rows_per_gpu": args.rows, "nsample": args.nsample, "n_fft": N_FFT, "hop": HOP,
            "global_rows": args.rows * world, "parallelism": f"utterance-sharded x{world}",
            "exchange": exchange,
            "l2": "working set ~400 MB/step > 126 MB L2; inputs rotate over 2 buffer sets"}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_chain(rows, nsample, steps, warmup, threads=None):
    """The reference's own CPU implementation of the step (oracle port: same torch entry points the
    reference calls, src/evaluate.py:101-162 + dcunet.py:131-161 + SURVEY 8c loss), timed on host cores."""
    import types
    import torch
    from oracle import spectral_oracle as oref
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it can
    threads = threads or len(os.sched_getaffinity(0)) or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = types.SimpleNamespace(n_fft=N_FFT, hop_length=HOP, win_length=WIN, center=True)
    g = torch.Generator().manual_seed(1235)
    x = torch.randn(rows, 1, nsample, generator=g)
    clean = x + 0.3 * torch.randn(rows, 1, nsample, generator=g)
    raw = torch.randn(rows, 1, N_FFT // 2 + 1, 1 + nsample // HOP, 2, generator=g).requires_grad_(True)

    def step():
        spec = oref.stft_custom_ref(x, cfg)
        y = oref.istft_custom_ref(oref.mask_apply_ref(spec, raw, "E", True), nsample, cfg)
        loss = oref.mrstft_loss_ref(y, clean)
        (graw,) = torch.autograd.grad(loss, raw)
        return float(loss.detach()), graw

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = min(args.cpu_rows, args.rows)
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 2))
    dt, threads = cpu_chain(rows, args.nsample, steps, warmup)
    value = rows * args.nsample / SR / dt
    sample = f"{rows} of {args.rows} rows per step, {steps} timed steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "host": host_info()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))



# ------------------------------------------------------------------------------------------ host placement
def numa_bind(local):
    """Bind this rank to the NUMA node of its GPU BEFORE any pinned allocation: CPU affinity to the node's cores (where
    the cpuset allows it) and a preferred-node memory policy, so that pinned staging buffers sit next to the GPU's PCIe
    root.  Round 1 measured 54 -> 23 GB/s per GPU at 8 ranks with every rank's buffers on node 0."""
    import ctypes
    import torch
    info = {"numa_node": None, "cpus": None, "mempolicy": None, "source": None}
    try:
        node, cpus = -1, set()
        try:
            pr = torch.cuda.get_device_properties(local)
            bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
            info["source"] = "sysfs"
        except Exception:
            node = -1
        if node < 0:
            # virtualised hosts report -1 in sysfs: ask NVML (what `nvidia-smi topo -m` prints)
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[local]) if visible and visible.split(",")[local].isdigit() else local
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64 + 1)
            for w, word in enumerate(words):
                cpus.update(64 * w + b for b in range(64) if (int(word) >> b) & 1)
            try:
                nodes = pynvml.nvmlDeviceGetMemoryAffinity(h, 4, 0)          # scope 0 = NUMA node
                node = next((64 * w + b for w, word in enumerate(nodes) for b in range(64) if (int(word) >> b) & 1), -1)
            except Exception:
                node = -1
            info["source"] = "nvml"
        info["numa_node"] = node
        if node >= 0 and not cpus:
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed and allowed != os.sched_getaffinity(0):
            os.sched_setaffinity(0, allowed)
        info["cpus"] = len(allowed)              # 0: the cpuset excludes that node's cores -> memory policy only
        if node >= 0:
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            rc = libc.syscall(238, 1, ctypes.byref(mask), 64)            # set_mempolicy(MPOL_PREFERRED, {node})  (x86_64)
            info["mempolicy"] = "preferred" if rc == 0 else f"errno {ctypes.get_errno()}"
    except Exception as e:                      # best effort: placement is an optimisation, not a requirement
        info["error"] = repr(e)[:160]
    return info


def host_info():
    import platform
    import torch
    model = platform.processor() or ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    mkl = ""
    try:
        mkl = next((l.strip(" -") for l in torch.__config__.show().splitlines() if "Math Kernel Library" in l), "")
    except Exception:
        pass
    return {"cpu_model": model, "logical_cpus": os.cpu_count(), "torch": torch.__version__, "mkl": mkl[:120]}


# ------------------------------------------------------------------------------------------ GPU incumbent
# The kernel set to beat: the reference's own torch calls on CUDA tensors -- torch.stft / torch.istft
# (cuFFT), ATen elementwise kernels, autograd -- restated here from src/evaluate.py:101-162, src/model/dcunet.py:131-155
# and the SURVEY 8c loss so that this leg does not touch oracle/.
def incumbent_step_fn(x, clean, raw, n_fft, hop, win):
    import torch

    def stft(t, n, h, w):
        r = torch.stft(t.reshape(-1, t.shape[-1]), n_fft=n, hop_length=h, win_length=w, window=torch.hann_window(w, device=t.device),
                       center=True, pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
        return torch.view_as_real(r)

    def step():
        raw.grad = None
        X = stft(x, n_fft, hop, win) / win
        m = torch.tanh(raw)
        xr, xi, mr, mi = X[..., 0], X[..., 1], m[..., 0], m[..., 1]
        x_mag, x_ph = torch.sqrt(xr ** 2 + xi ** 2 + 1e-8), torch.atan2(xi, xr)
        m_mag = (mr ** 2 + mi ** 2) ** 0.5
        m_ph = torch.atan2(mi / (m_mag + 1e-8), mr / (m_mag + 1e-8))
        e_mag, e_ph = torch.tanh(m_mag) * x_mag, x_ph + m_ph
        Y = torch.complex(e_mag * torch.cos(e_ph), e_mag * torch.sin(e_ph)) * win
        y = torch.istft(Y, n_fft=n_fft, hop_length=hop, win_length=win, window=torch.hann_window(win, device=x.device),
                        center=True, normalized=False, onesided=True, length=x.shape[-1], return_complex=False)
        total = 0.0
        for n, h in RES:
            A, B = stft(y, n, h, n), stft(clean, n, h, n)
            a = torch.sqrt(torch.clamp(A[..., 0] ** 2 + A[..., 1] ** 2, min=1e-7))
            b = torch.sqrt(torch.clamp(B[..., 0] ** 2 + B[..., 1] ** 2, min=1e-7))
            total = total + torch.linalg.norm((b - a).reshape(-1)) / torch.linalg.norm(b.reshape(-1)) + torch.mean(torch.abs(torch.log(b) - torch.log(a)))
        (total / len(RES)).backward()
        return total
    return step


# ------------------------------------------------------------------------------------------ the other BASELINE configs
def run_configs(se, dev, world, group, hbm_peak, fp32_peak):
    """BASELINE.json configs 1, 3, 4, 5 through the public API (the calls a user makes), device-resident inputs rotated
    over 2-4 buffer sets, CUDA events, max over ranks; every rank runs the per-GPU share (weak scaling, no collective
    except cfg3's 9-double exchange).  cfg2 is the headline above."""
    import types
    import torch
    import torch.distributed as dist

    def cfg(n, h):
        return types.SimpleNamespace(n_fft=n, hop_length=h, win_length=n, center=True)

    def timed(fn, sets, reps=30, warm=5):
        for i in range(warm):
            fn(sets[i % len(sets)])
        torch.cuda.synchronize(dev)
        if group is not None:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            fn(sets[i % len(sets)])
        b.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([a.elapsed_time(b) / reps * 1e3], device=dev)
        if group is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    out = []

    def entry(name, workload, us, audio_s_per_gpu, alg_bytes=None, flops=None, **extra):
        e = {"name": name, "workload": workload, "us": round(us, 1), "value": round(audio_s_per_gpu * world / us * 1e6), "unit": UNIT,
             "n_gpus": world}
        if alg_bytes:
            e["alg_bytes_per_gpu"] = int(alg_bytes)
            e["hbm_frac"] = round(alg_bytes / us * 1e-3 / hbm_peak, 4)
        if flops:
            e["fp32_frac"] = round(flops / us * 1e-6 / fp32_peak, 4)
        e.update(extra)
        out.append(e)

    g = torch.Generator().manual_seed(1234)
    N = 64000
    # ---- cfg1: 16 x 4 s, n_fft 512 / hop 128, real (Unet) mask, inference
    c = cfg(512, 128)
    sets = [(torch.randn(16, 1, N, generator=g).to(dev), torch.rand(16, 1, 257, 501, generator=g).to(dev)) for _ in range(4)]
    S, P, M = 4 * N, 8 * 257 * 501, 4 * 257 * 501
    with torch.no_grad():
        us = timed(lambda s: se.istft_custom(se.apply_mask(se.stft_custom(s[0], c), s[1], "real"), N, c), sets, reps=50)
        entry("cfg1_dropin", "16x4 s, 512/128, stft_custom -> apply_mask(real) -> istft_custom (3 API calls), no-grad", us, 64,
              16 * (2 * S + 4 * P + M))
        us = timed(lambda s: se.enhance(s[0], s[1], c, "real"), sets, reps=50)
        entry("cfg1_fused", "16x4 s, 512/128, se.enhance (one launch), no-grad", us, 64, 16 * (2 * S + M))
    # ---- cfg3: MR-STFT loss fwd+bwd, 128 x 4 s per GPU
    sets = [(torch.randn(128, 1, N, generator=g).to(dev).requires_grad_(True), torch.randn(128, 1, N, generator=g).to(dev)) for _ in range(2)]

    def loss_step(s):
        s[0].grad = None
        se.loss_mrstft(s[0], s[1], group).backward()
    us = timed(loss_step, sets, reps=20)
    flops = 128 * 4 * sum(2.5 * n * (n.bit_length() - 1) * (1 + N // h) for n, h in RES)      # 4 transforms per resolution
    entry("cfg3_mrstft_loss", "MR-STFT loss 512/1024/2048 fwd+bwd, 128x4 s per GPU (loss_mrstft + backward)", us, 512,
          128 * 5 * S, flops, bound="fp32 pipe / L1TEX (75 flop/B)")
    del sets
    # ---- cfg4: DCCRN in-model transforms, 16 x 4 s per GPU, fp32 spectra: ConvSTFT -> mask tail + ConviSTFT, fwd + bwd to the masks
    st, ist = se.ConvSTFT(400, 100, 512, "hann", "complex"), se.ConviSTFT(400, 100, 512, N, "hann", "complex")
    sets = [(torch.randn(16, 1, N, generator=g).to(dev), torch.randn(16, 257, 643, generator=g).to(dev).requires_grad_(True),
             torch.randn(16, 257, 643, generator=g).to(dev).requires_grad_(True)) for _ in range(4)]

    def dccrn_step(s):
        s[1].grad = None
        s[2].grad = None
        y = ist.forward_masked(st(s[0]), s[1], s[2], "E")
        y.backward(y)                                           # stand-in upstream gradient
    us = timed(dccrn_step, sets, reps=30)
    Pp = 4 * 514 * 643
    entry("cfg4_dccrn_transforms", "DCCRN ConvSTFT -> 'E' mask tail + ConviSTFT (forward_masked), fwd+bwd to the masks, 16x4 s per GPU",
          us, 64, 16 * ((S + Pp) + (Pp + Pp + S) + (S + Pp + 2 * Pp)))
    del sets
    # ---- cfg5: 44.1 kHz stereo 30 s clips, 8 clips per GPU, complex mask, inference; n_fft 2048 and 1024
    NL = 1323000
    for n in (2048, 1024):
        c = cfg(n, n // 4)
        F, T = n // 2 + 1, 1 + NL // (n // 4)
        sets = [(torch.randn(8, 2, NL, generator=g).to(dev), (torch.rand(8, 2, F, T, 2, generator=g) * 2 - 1).to(dev)) for _ in range(2)]
        with torch.no_grad():
            us = timed(lambda s: se.enhance(s[0], s[1], c, "C"), sets, reps=10, warm=3)
        entry(f"cfg5_enhance_n{n}", f"8 clips x 30 s stereo 44.1 kHz per GPU, n_fft {n} hop {n // 4}, se.enhance complex mask, inference",
              us, 240, 16 * (2 * 4 * NL + 8 * F * T), channel_s_per_s=round(480 * world / us * 1e6))
        del sets
        torch.cuda.empty_cache()
    # ---- cfg5, reference-faithful variant: the reference resamples to 16 kHz and runs evaluate() -- z-score, 4 s
    # segments at stride win_length = 512 (814 segments of a 30 s clip), STFT, [model], iSTFT, stitch; one stereo clip per GPU
    conf = types.SimpleNamespace(dset=types.SimpleNamespace(norm="z-score", sample_rate=16000),
                                 model=types.SimpleNamespace(name="unet", segment=4.0, n_fft=512, hop_length=128, win_length=512,
                                                             center=True, sources=["clean"]))
    sets = [((0.3 * torch.randn(1, 2, 480000, generator=g) + 0.01).to(dev),) for _ in range(2)]
    us = timed(lambda s: se.evaluate(s[0], None, dev, conf), sets, reps=10, warm=3)
    entry("cfg5_evaluate_16k", "se.evaluate(model=None): one 30 s stereo clip at 16 kHz per GPU, 814 segments x 4 s at stride 512 "
          "(row stats, shared-frame segment STFT, stitching iSTFT)", us, 30, 2 * 4 * 480000 * 2 + 8 * 257 * 501 * 814 * 2,
          channel_s_per_s=round(60 * world / us * 1e6))
    del sets
    # ---- general-geometry tier (csrc/se_generic.cuh): geometries outside the reference's configs, drop-in transforms only.
    # Reported for completeness; nothing else in this line is measured on it.
    for n, h in ((1024, 255), (320, 160)):
        c = cfg(n, h)
        F, T = n // 2 + 1, 1 + N // h
        sets = [(torch.randn(64, 1, N, generator=g).to(dev),) for _ in range(2)]
        with torch.no_grad():
            us = timed(lambda s: se.istft_custom(se.stft_custom(s[0], c), N, c), sets, reps=10, warm=3)
        entry(f"general_{n}_{h}", f"general-geometry tier: 64x4 s, n_fft {n} hop {h}, stft_custom -> istft_custom, no-grad", us, 256,
              64 * (2 * S + 2 * 8 * F * T))
        del sets
    return out


def cpu_configs_parity(se, dev):
    """cpu_baseline leg only: the oracle (the reference's CPU entry points) on a bounded sample of every config, timed on
    the host cores and used as the checker of the GPU result on the same inputs."""
    import types
    import torch
    from oracle import spectral_oracle as oref
    res = []

    def rel(a, b):
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

    def rel_l2(a, b):
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        return float((a - b).norm() / b.norm().clamp_min(1e-30))

    def cfg(n, h):
        return types.SimpleNamespace(n_fft=n, hop_length=h, win_length=n, center=True)

    g = torch.Generator().manual_seed(1234)
    N = 64000
    # cfg1: 4 of 16 rows
    c = cfg(512, 128)
    x, m = torch.randn(4, 1, N, generator=g), torch.rand(4, 1, 257, 501, generator=g)
    t0 = time.perf_counter()
    want = oref.istft_custom_ref(oref.mask_apply_ref(oref.stft_custom_ref(x, c), m, "real"), N, c)
    dt = time.perf_counter() - t0
    with torch.no_grad():
        got = se.enhance(x.to(dev), m.to(dev), c, "real")
    res.append({"name": "cfg1", "cpu_audio_s_per_s": round(16 / dt), "sample": "4 of 16 rows", "max_rel_err": rel(got, want), "tol": 1e-4})
    # cfg3: 4 of 128 rows, loss + gradient (float64 oracle is the ground truth for the gradient)
    ref = torch.randn(4, 1, N, generator=g)
    est = ref + 0.1 * torch.randn(4, 1, N, generator=g)
    e32 = est.clone().requires_grad_(True)
    t0 = time.perf_counter()
    l32 = oref.mrstft_loss_ref(e32, ref)
    (g32,) = torch.autograd.grad(l32, e32)
    dt = time.perf_counter() - t0
    e64 = est.double().requires_grad_(True)
    l64 = oref.mrstft_loss_ref(e64, ref.double())
    (g64,) = torch.autograd.grad(l64, e64)
    eg = est.to(dev).requires_grad_(True)
    lg = se.loss_mrstft(eg, ref.to(dev))
    (gg,) = torch.autograd.grad(lg, eg)
    res.append({"name": "cfg3", "cpu_audio_s_per_s": round(16 / dt), "sample": "4 of 128 rows, fwd+bwd",
                "loss_rel_err": abs(float(lg) - float(l64)) / float(l64), "grad_rel_l2_vs_f64": rel_l2(gg, g64),
                "grad_max_rel_vs_f64": rel(gg, g64), "reference_fp32_grad_rel_l2_vs_f64": rel_l2(g32, g64), "tol": 1e-3})
    # cfg4: 2 of 16 rows, DCCRN transforms
    x = torch.randn(2, 1, N, generator=g)
    mre, mim = torch.randn(2, 257, 643, generator=g), torch.randn(2, 257, 643, generator=g)
    t0 = time.perf_counter()
    spec = oref.conv_stft_ref(x, 400, 100, 512)
    re_, im_ = spec[:, :257], spec[:, 257:]
    masked = oref.mask_apply_ref(torch.stack([re_, im_], -1), torch.stack([mre, mim], -1), "E")
    want = oref.conv_istft_ref(torch.cat([masked[..., 0], masked[..., 1]], 1), 400, 100, 512, length=N)
    dt = time.perf_counter() - t0
    st, ist = se.ConvSTFT(400, 100, 512, "hann", "complex"), se.ConviSTFT(400, 100, 512, N, "hann", "complex")
    with torch.no_grad():
        got = ist.forward_masked(st(x.to(dev)), mre.to(dev), mim.to(dev), "E")
    res.append({"name": "cfg4", "cpu_audio_s_per_s": round(8 / dt), "sample": "2 of 16 rows, forward", "max_rel_err": rel(got.reshape(want.shape), want),
                "tol": 1e-4})
    # cfg5: one stereo clip
    NL = 1323000
    for n in (2048, 1024):
        c = cfg(n, n // 4)
        F, T = n // 2 + 1, 1 + NL // (n // 4)
        x, m = torch.randn(1, 2, NL, generator=g), torch.rand(1, 2, F, T, 2, generator=g) * 2 - 1
        t0 = time.perf_counter()
        want = oref.istft_custom_ref(oref.mask_apply_ref(oref.stft_custom_ref(x, c), m, "C"), NL, c)
        dt = time.perf_counter() - t0
        with torch.no_grad():
            got = se.enhance(x.to(dev), m.to(dev), c, "C")
        res.append({"name": f"cfg5_n{n}", "cpu_audio_s_per_s": round(30 / dt), "sample": "1 of 8 clips", "max_rel_err": rel(got, want), "tol": 1e-4})
    # cfg5 through evaluate(): a 6 s stereo clip (150 segments) against the oracle's restatement of src/evaluate.py:10-98
    conf = types.SimpleNamespace(dset=types.SimpleNamespace(norm="z-score", sample_rate=16000),
                                 model=types.SimpleNamespace(name="unet", segment=4.0, n_fft=512, hop_length=128, win_length=512,
                                                             center=True, sources=["clean"]))
    x = 0.3 * torch.randn(1, 2, 96000, generator=g) + 0.01
    t0 = time.perf_counter()
    want = oref.evaluate_ref(x, None, conf)
    dt = time.perf_counter() - t0
    got = se.evaluate(x, None, dev, conf)
    res.append({"name": "cfg5_evaluate_16k", "cpu_audio_s_per_s": round(6 / dt, 1), "sample": "one 6 s stereo clip (63 segments)",
                "max_rel_err": rel(got, want), "tol": 1e-4})
    return res


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [v.strip() for v in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    import speech_enhancement_pytorch_b200 as se
    from speech_enhancement_pytorch_b200 import _native as nv
    from speech_enhancement_pytorch_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    placement = numa_bind(local)            # before any pinned allocation (and before NCCL spawns its threads)
    group = None
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: park stdout on
        # stderr for the run so that rank 0's stdout carries exactly one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    L = nv.lib()
    # the path's one exchange step: fused peer-memory kernel (exchange + loss value) when the ranks can map each
    # other's buffers, else NCCL all-reduce followed by the value kernel
    px = None
    if group is not None:
        from speech_enhancement_pytorch_b200 import distributed as sed
        px = sed.peer_exchange(group, dev)
    exchange = ("none (1 GPU)" if group is None else
                "9 doubles, one fused peer-memory kernel over NVLink (se_mrstft_exchange_value)" if px is not None else
                "9 doubles, NCCL all-reduce on the compute stream")
    rows, N = args.rows, args.nsample
    F, T = N_FFT // 2 + 1, 1 + N // HOP
    # three compositions of the same step, all timed and reported; the first is the default:
    #   "tail"   stft_custom, then apply_mask_istft (mask tail + iSTFT in one launch each way; the masked
    #            spectrum and its gradient are never written) -- 3 launches around the loss
    #   "dropin" stft_custom / apply_mask / istft_custom one kernel each -- 5 launches
    #   "fused"  se.enhance: everything in one launch each way (the backward recomputes the STFT) -- 2 launches
    comp = {0: "dropin", 1: "fused"}.get(args.fused, "tail")

    # ---- device-resident inputs (2 rotating sets) and preallocated intermediates
    g = torch.Generator(device="cpu").manual_seed(1235 + rank)
    sets = []
    for _ in range(2):
        x = torch.randn(rows, N, generator=g)
        clean = x + 0.3 * torch.randn(rows, N, generator=g)
        raw = torch.randn(rows, F, T, 2, generator=g)
        sets.append((x.to(dev), clean.to(dev), raw.to(dev)))
    X = torch.empty(rows, F, T, 2, device=dev)
    Y = torch.empty_like(X)
    y = torch.empty(rows, N, device=dev)
    gy = torch.empty(rows, N, device=dev)
    gY = torch.empty_like(X)
    graw = torch.empty_like(X)
    ws = torch.empty(max(int(L.se_mrstft_workspace_bytes(rows, N)), 8), dtype=torch.uint8, device=dev)
    sums = torch.empty(9, dtype=torch.float64, device=dev)      # raw C-ABI step: equal shards, host-known global row count
    loss = torch.empty((), device=dev)
    one = torch.ones((), device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    P = lambda t: t.data_ptr()
    count = rows * F * T
    launches = {"n": 0}

    def k_stft(x): nv.check(L.se_stft_fwd(P(x), P(X), rows, N, N_FFT, HOP, WIN, 1.0 / WIN, st))
    def k_mask(raw): nv.check(L.se_mask_fwd(P(X), P(raw), P(Y), count, 1, 1, st))
    def k_istft(): nv.check(L.se_istft_fwd(P(Y), P(y), rows, T, N, N_FFT, HOP, WIN, float(WIN), st))
    def k_loss_fwd(clean): nv.check(L.se_mrstft_loss_fwd(P(y), P(clean), rows, N, P(sums), P(ws), st))
    def k_loss_fwd_value(clean): nv.check(L.se_mrstft_loss_fwd_value(P(y), P(clean), rows, N, P(sums), P(loss), P(ws), st))
    def k_loss_val(): nv.check(L.se_mrstft_loss_value(P(sums), rows * world, N, P(loss), st))
    def k_loss_bwd(clean): nv.check(L.se_mrstft_loss_bwd(P(y), P(ws), P(sums), P(one), rows * world, rows, N, P(gy), st))
    def k_istft_bwd(): nv.check(L.se_istft_bwd(P(gy), P(gY), rows, T, N, N_FFT, HOP, WIN, float(WIN), st))
    def k_mask_bwd(raw): nv.check(L.se_mask_bwd(P(X), P(raw), P(gY), P(graw), 0, count, 1, 1, st))
    def k_enh_fwd(x, raw): nv.check(L.se_enhance_fwd(P(x), P(raw), P(y), rows, N, N_FFT, HOP, WIN, 1, 1, st))
    def k_enh_bwd(x, raw): nv.check(L.se_enhance_bwd(P(gy), P(x), P(raw), P(graw), rows, N, N_FFT, HOP, WIN, 1, 1, st))
    def k_tail_fwd(raw): nv.check(L.se_mask_istft_fwd(P(X), P(raw), P(y), rows, T, N, N_FFT, HOP, WIN, float(WIN), 1, 1, st))
    def k_tail_bwd(raw): nv.check(L.se_mask_istft_bwd(P(gy), P(X), P(raw), P(graw), rows, T, N, N_FFT, HOP, WIN, float(WIN), 1, 1, st))

    def step(i, comp=comp):
        x, clean, raw = sets[i & 1]
        if comp == "fused":
            k_enh_fwd(x, raw)
        elif comp == "tail":
            k_stft(x); k_tail_fwd(raw)
        else:
            k_stft(x); k_mask(raw); k_istft()
        if group is None:
            k_loss_fwd_value(clean)                              # one process: the reduction launch also writes the loss value
        else:
            k_loss_fwd(clean)
            if px is not None:
                px.exchange_value(sums, rows * world, N, loss, st)   # the path's only exchange step, fused with the value
            else:
                dist.all_reduce(sums, group=group)
                k_loss_val()
        k_loss_bwd(clean)
        if comp == "fused":
            k_enh_bwd(x, raw)
        elif comp == "tail":
            k_tail_bwd(raw)
        else:
            k_istft_bwd(); k_mask_bwd(raw)

    # + 3 loss fwd, 1 reduce (which writes the value when there is one process; else + 1 exchange / value launch), 3 loss bwd
    n_launch = {"fused": 2, "tail": 3, "dropin": 5}[comp] + 4 + (0 if group is None else 1) + 3

    def sync_all():
        torch.cuda.synchronize(dev)
        if group is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        step(i)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # `rounds` timed rounds of EXACTLY --steps steps, each bracketed by barrier + synchronize on both sides and reduced
    # with max over ranks; the median round is the reported one (median, not a single shot)
    rounds_ms = []
    for _ in range(max(args.rounds, 1)):
        sync_all()
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if group is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        rounds_ms.append(float(ms) / args.steps)
    ms_per_step = sorted(rounds_ms)[len(rounds_ms) // 2]
    # ---- N > 1: the sharded loss (9-double exchange step) must equal the single-GPU loss over the concatenated batch.
    # Every rank's first input set is regenerated from its seed on rank 0 (the driver's box has the GPUs the 2-rank
    # pytest cases need, but runs them on one): printed in the line, and the run fails if they disagree.
    shard_check = None
    if group is not None:
        xs, cs = sets[0][0], sets[0][1]
        est = xs.reshape(rows, 1, N).clone().requires_grad_(True)
        l_sh = se.loss_mrstft(est, cs.reshape(rows, 1, N), group)
        (g_sh,) = torch.autograd.grad(l_sh, est)
        if rank == 0:
            allx, allc = [], []
            for r in range(world):
                gr = torch.Generator(device="cpu").manual_seed(1235 + r)
                xr = torch.randn(rows, N, generator=gr)
                allx.append(xr)
                allc.append(xr + 0.3 * torch.randn(rows, N, generator=gr))
            est_all = torch.cat(allx).reshape(rows * world, 1, N).to(dev).requires_grad_(True)
            l_1 = se.loss_mrstft(est_all, torch.cat(allc).reshape(rows * world, 1, N).to(dev))
            (g_1,) = torch.autograd.grad(l_1, est_all)
            gd = float((g_sh - g_1[:rows]).abs().max() / g_1[:rows].abs().max())
            shard_check = {"loss_sharded": float(l_sh), "loss_single_gpu": float(l_1),
                           "rel_diff": abs(float(l_sh) - float(l_1)) / abs(float(l_1)), "grad_rel_diff_rank0_rows": gd,
                           "what": f"loss_mrstft over {world} ranks x {rows} rows vs one GPU on all {rows * world} rows, same inputs"}
            assert shard_check["rel_diff"] < 1e-6 and gd < 1e-5, shard_check
            del est_all, g_1, allx, allc
        del est, g_sh
        torch.cuda.empty_cache()
        sync_all()

    # the other compositions, same timing rules, for context
    audio_s = rows * world * N / SR
    alts = []
    for other in ("tail", "dropin", "fused"):
        if other == comp:
            continue
        for i in range(3):
            step(i, other)
        sync_all()
        e0.record()
        for i in range(args.steps):
            step(i, other)
        e1.record()
        sync_all()
        ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if group is not None:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        alt_ms = float(ms2) / args.steps
        alts.append({"composition": other, "ms_per_step": alt_ms, "value": audio_s / (alt_ms * 1e-3)})
    value = audio_s / (ms_per_step * 1e-3)
    loss_val = float(loss)

    # ---- per-kernel breakdown (CUDA events on the launching stream), rank 0 only
    kernels = []
    if rank == 0 and not args.no_breakdown:
        S_, P_, M_ = 4.0 * N, 8.0 * F * T, 8.0 * F * T
        # algorithmic bytes per row follow SURVEY.md 8(d): op-boundary traffic, independent of the number of
        # resolutions (loss fwd 2S, loss bwd 3S); the |B| scratch the two passes exchange is implementation traffic
        flop_fft = lambda n, h: 2.5 * n * (n.bit_length() - 1) * (1 + N // h)
        f1024 = flop_fft(N_FFT, HOP)
        f_all = sum(flop_fft(n, h) for n, h in RES)
        table = [
            ("enhance_fwd", lambda i: k_enh_fwd(sets[i & 1][0], sets[i & 1][2]), 2 * S_ + M_, 2 * f1024, 1),
            ("enhance_bwd", lambda i: k_enh_bwd(sets[i & 1][0], sets[i & 1][2]), 2 * S_ + 2 * M_, 2 * f1024, 1),
            ("mask_istft_fwd", lambda i: k_tail_fwd(sets[i & 1][2]), P_ + M_ + S_, f1024, 1),
            ("mask_istft_bwd", lambda i: k_tail_bwd(sets[i & 1][2]), S_ + P_ + 2 * M_, f1024, 1),
            ("stft_fwd", lambda i: k_stft(sets[i & 1][0]), S_ + P_, f1024, 1),
            ("mask_fwd", lambda i: k_mask(sets[i & 1][2]), 2 * P_ + M_, 0, 1),
            ("istft_fwd", lambda i: k_istft(), P_ + S_, f1024, 1),
            ("mrstft_loss_fwd(3 res)", lambda i: k_loss_fwd(sets[i & 1][1]), 2 * S_, 2 * f_all, 4),
            ("mrstft_loss_bwd(3 res)", lambda i: k_loss_bwd(sets[i & 1][1]), 3 * S_, 2 * f_all, 3),
            ("istft_bwd", lambda i: k_istft_bwd(), S_ + P_, f1024, 1),
            ("mask_bwd", lambda i: k_mask_bwd(sets[i & 1][2]), 2 * P_ + 2 * M_, 0, 1),
        ]
        reps = 20
        for name, fn, bytes_row, flops_row, nl in table:
            for i in range(3):
                fn(i)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(reps):
                fn(i)
            b.record()
            torch.cuda.synchronize(dev)
            us = a.elapsed_time(b) * 1e3 / reps
            kernels.append({"name": name, "us": round(us, 2), "launches": nl,
                            "alg_bytes": bytes_row * rows, "gbs": round(bytes_row * rows / us * 1e-3, 1),
                            "tflops_fp32": round(flops_row * rows / us * 1e-6, 2)})

    # ---- what the reference-facing API adds on top of the raw C-ABI launch (no-grad calls, same shapes)
    api_overhead = None
    if rank == 0 and not args.no_breakdown:
        import types as _types
        cfg0 = _types.SimpleNamespace(n_fft=N_FFT, hop_length=HOP, win_length=WIN, center=True)
        x5 = sets[0][0].reshape(rows, 1, N)
        with torch.no_grad():
            spec5 = se.stft_custom(x5, cfg0)

            def t_api(fn, reps=50):
                for _ in range(5):
                    fn()
                torch.cuda.synchronize(dev)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    fn()
                b.record()
                torch.cuda.synchronize(dev)
                return a.elapsed_time(b) * 1e3 / reps
            raw_us = {k["name"]: k["us"] for k in kernels}
            api_overhead = {"stft_custom_us": round(t_api(lambda: se.stft_custom(x5, cfg0)), 2), "se_stft_fwd_us": raw_us.get("stft_fwd"),
                            "istft_custom_us": round(t_api(lambda: se.istft_custom(spec5, N, cfg0)), 2), "se_istft_fwd_us": raw_us.get("istft_fwd"),
                            "note": "public API (C++ extension op: checks, output allocation, current stream) vs the raw C-ABI launch with "
                                    "preallocated outputs, back-to-back calls, CUDA events"}
        del spec5

    # the sampler ran through the timed region, the alternative composition and the per-kernel loops (all under load)
    clocks = sampler.stop() if rank == 0 else None

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    fp32_peak = 148 * 128 * 2 * (float(clocks["sm_max_mhz"]) if clocks and clocks.get("sm_max_mhz") else 1965.0) * 1e-6

    # ---- the GPU incumbent of the same step: torch.stft / torch.istft (cuFFT) + ATen + autograd, rank 0 only
    incumbent = None
    if rank == 0 and not args.no_incumbent:
        x0, c0, r0 = sets[0]
        rawp = r0.clone().requires_grad_(True)                       # [rows, F, T, 2]
        inc = incumbent_step_fn(x0, c0, rawp, N_FFT, HOP, WIN)        # waveforms [rows, N]
        for _ in range(3):
            inc()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            inc()
        b.record()
        torch.cuda.synchronize(dev)
        inc_ms = a.elapsed_time(b) / 10
        incumbent = {"what": "the reference's own torch calls on CUDA tensors: torch.stft/istft (cuFFT) + ATen elementwise + autograd, "
                             "same cfg2 step, same GPU, 10 timed steps", "ms_per_step": inc_ms, "value": rows * N / SR / (inc_ms * 1e-3),
                     "unit": UNIT, "speedup_ours": inc_ms / ms_per_step}
        del inc, rawp
        torch.cuda.empty_cache()

    # ---- BASELINE configs 1, 3, 4, 5 in the same run (all ranks: per-GPU share each)
    configs = None
    if not args.no_configs:
        sync_all()
        configs = run_configs(se, dev, world, group, hbm_peak, fp32_peak)

    # ---- end-to-end through the public API with host buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        import types
        sync_all()          # rank 0 alone ran the per-kernel loops: line the ranks up before the next exchange step
        cfg = types.SimpleNamespace(n_fft=N_FFT, hop_length=HOP, win_length=WIN, center=True)
        hx = [torch.randn(rows, 1, N, generator=g).pin_memory() for _ in range(2)]
        hc = [(hx[i] + 0.3 * torch.randn(rows, 1, N, generator=g)).pin_memory() for i in range(2)]
        hm = [torch.randn(rows, 1, F, T, 2, generator=g).pin_memory() for _ in range(2)]
        hloss = torch.empty((), pin_memory=True)
        copy_stream = torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)
        dbuf = [(torch.empty(rows, 1, N, device=dev), torch.empty(rows, 1, N, device=dev),
                 torch.empty(rows, 1, F, T, 2, device=dev)) for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def upload(i):
            j = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[j])
                dbuf[j][0].copy_(hx[j], non_blocking=True)
                dbuf[j][1].copy_(hc[j], non_blocking=True)
                dbuf[j][2].copy_(hm[j], non_blocking=True)
                ready[j].record(copy_stream)

        def e2e_step(i):
            j = i & 1
            main.wait_event(ready[j])
            x, clean = dbuf[j][0], dbuf[j][1]
            raw = dbuf[j][2].detach().requires_grad_(True)
            if comp == "fused":
                yy = se.enhance(x, raw, cfg, "E", True)
            elif comp == "tail":
                yy = se.apply_mask_istft(se.stft_custom(x, cfg), raw, N, cfg, "E", True)
            else:
                yy = se.istft_custom(se.apply_mask(se.stft_custom(x, cfg), raw, "E", True), N, cfg)
            l = se.loss_mrstft(yy, clean, group)
            l.backward()
            hloss.copy_(l.detach(), non_blocking=True)
            consumed[j].record(main)
            return raw.grad

        # the host's own ceiling: the same three pinned uploads back to back with no compute (all ranks at once) -- what
        # the PCIe / host-memory path of this box can deliver per GPU at this N
        sync_all()
        ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(copy_stream):
            ca.record(copy_stream)
            for i in range(10):
                for d, h in zip(dbuf[i & 1], (hx[i & 1], hc[i & 1], hm[i & 1])):
                    d.copy_(h, non_blocking=True)
            cb.record(copy_stream)
        sync_all()
        copy_only = torch.tensor([(hx[0].numel() + hc[0].numel() + hm[0].numel()) * 4 * 10 / (ca.elapsed_time(cb) * 1e-3) * 1e-9], device=dev)
        if group is not None:
            dist.all_reduce(copy_only, op=dist.ReduceOp.MIN)
        for j in range(2):
            consumed[j].record(main)
        e2e_steps = max(5, min(args.steps, 30))
        upload(0)
        for i in range(3):
            upload(i + 1)
            e2e_step(i)
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        upload(3)
        a.record()
        for i in range(3, 3 + e2e_steps):
            upload(i + 1)
            e2e_step(i)
        b.record()
        sync_all()
        t = torch.tensor([a.elapsed_time(b)], device=dev)
        if group is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        own_ms = float(a.elapsed_time(b)) / e2e_steps
        e2e_ms = float(t) / e2e_steps
        h2d = (hx[0].numel() + hc[0].numel() + hm[0].numel()) * 4
        per_rank = torch.tensor([h2d / (own_ms * 1e-3) * 1e-9, float(placement.get("numa_node") if placement.get("numa_node") is not None else -1)],
                                device=dev)
        if group is not None:
            gathered = [torch.zeros_like(per_rank) for _ in range(world)]
            dist.all_gather(gathered, per_rank)
        else:
            gathered = [per_rank]
        e2e = {"value": audio_s / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": e2e_steps,
               "h2d_gbs": round(h2d / (e2e_ms * 1e-3) * 1e-9, 1),
               "h2d_gbs_per_rank": [round(float(v[0]), 1) for v in gathered],
               "numa_node_per_rank": [int(v[1]) for v in gathered], "placement_rank0": placement,
               "h2d_copy_only_gbs_min_rank": round(float(copy_only), 1),
               "note": "copy-bound: the 98.7 MB/step of pinned-host inputs (mixture, clean, raw mask) saturate the host-to-device path; "
                       "kernels overlap underneath on the compute stream.  h2d_copy_only_gbs_min_rank is the same uploads with no "
                       "compute at this N: the pool's 8-GPU boxes are VMs with ONE NUMA node (nvidia-smi topo: all GPUs NUMA 0, "
                       "cpus 0-31), so per-rank placement cannot help and the aggregate host ceiling (~185 GB/s) bounds e2e scaling",
               "api": {"tail": "stft_custom/apply_mask_istft", "dropin": "stft_custom/apply_mask/istft_custom",
                       "fused": "enhance"}[comp] + "/loss_mrstft + autograd; pinned host inputs, copy stream double-buffered",
               "loss": float(hloss)}

        # Supplementary (not the headline): what a training loop sees, where the raw mask is produced ON the device by
        # the model.  The NN bodies are out of scope, so a pointwise stand-in makes the raw mask from the
        # spectrum; host inputs are the two waveforms only (what the reference's DataLoader hands the solver).
        def upload_wav(i):
            j = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[j])
                dbuf[j][0].copy_(hx[j], non_blocking=True)
                dbuf[j][1].copy_(hc[j], non_blocking=True)
                ready[j].record(copy_stream)

        def e2e_step_model(i):
            j = i & 1
            main.wait_event(ready[j])
            x, clean = dbuf[j][0], dbuf[j][1]
            spec = se.stft_custom(x, cfg)
            # stand-in for the NN body: the identity (raw mask = the device spectrum itself, no kernel), so that this entry
            # isolates the API's own cost over the device-timed `value`; a real model adds its own time here
            raw = spec.detach().requires_grad_(True)
            yy = se.apply_mask_istft(spec, raw, N, cfg, "E", True)
            l = se.loss_mrstft(yy, clean, group)
            l.backward()
            hloss.copy_(l.detach(), non_blocking=True)
            consumed[j].record(main)

        sync_all()
        for j in range(2):
            consumed[j].record(main)
        upload_wav(0)
        for i in range(3):
            upload_wav(i + 1)
            e2e_step_model(i)
        sync_all()
        upload_wav(3)
        a.record()
        for i in range(3, 3 + e2e_steps):
            upload_wav(i + 1)
            e2e_step_model(i)
        b.record()
        sync_all()
        t = torch.tensor([a.elapsed_time(b)], device=dev)
        if group is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        m_ms = float(t) / e2e_steps
        # the same API step with the waveforms already on the device: isolates what the public API (C++ extension ops,
        # autograd, allocator) adds over the raw C-ABI step timed as `value` -- no copies in this one
        def api_step_device(i):
            x, clean = dbuf[i & 1][0], dbuf[i & 1][1]
            spec = se.stft_custom(x, cfg)
            raw = spec.detach().requires_grad_(True)
            l = se.loss_mrstft(se.apply_mask_istft(spec, raw, N, cfg, "E", True), clean, group)
            l.backward()
        sync_all()
        for i in range(3):
            api_step_device(i)
        sync_all()
        a.record()
        for i in range(e2e_steps):
            api_step_device(i)
        b.record()
        sync_all()
        t = torch.tensor([a.elapsed_time(b)], device=dev)
        if group is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        api_ms = float(t) / e2e_steps
        e2e["api_step_device_inputs"] = {"ms_per_step": api_ms, "vs_device_timed_value": round(api_ms / ms_per_step, 4),
                                         "note": "public API + autograd on device-resident waveforms (no copies): API cost over the raw C-ABI step"}
        e2e["mask_made_on_device"] = {
            "value": audio_s / (m_ms * 1e-3), "unit": UNIT, "ms_per_step": m_ms,
            "h2d_bytes_per_step": (hx[0].numel() + hc[0].numel()) * 4, "d2h_bytes_per_step": 4,
            "note": "supplementary: raw mask = identity stand-in model on the device spectrum (no kernel); host inputs are mixture + clean only "
                    "(32.8 MB per step: at ~54 GB/s that upload alone is 0.61 ms, so this entry is copy-bound too)",
            "vs_device_timed_value": round(m_ms / ms_per_step, 4)}

    if rank != 0:
        if group is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    members = {"fused": ("enhance_fwd", "enhance_bwd"), "tail": ("stft_fwd", "mask_istft_fwd", "mask_istft_bwd"),
               "dropin": ("stft_fwd", "mask_fwd", "istft_fwd", "istft_bwd", "mask_bwd")}[comp]
    in_step = [k for k in kernels if k["name"].startswith("mrstft") or k["name"] in members]
    dom = max(in_step, key=lambda k: k["us"]) if in_step else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom["name"])
    except Exception:
        pass
    roofline = None
    if dom:
        per_launch_bytes = dom["alg_bytes"] / dom["launches"]
        per_launch_us = dom["us"] / dom["launches"]
        roofline = {"kernel": dom["name"], "bound": "hbm",
                    "limiter": "not HBM: shared-memory (L1TEX) wavefronts + fp32 issue at ~15 warps/SM (ncu, profiles/r02_notes.md); "
                               "the hbm fraction is the contract's figure, the fp32 block the informative one", "achieved": round(per_launch_bytes / per_launch_us * 1e-3, 1),
                    "peak": hbm_peak, "unit": "GB/s", "frac": round(per_launch_bytes / per_launch_us * 1e-3 / hbm_peak, 4),
                    "traffic": traffic, "peak_source": peak_src,
                    "share_of_step": round(dom["us"] / sum(k["us"] for k in in_step), 3),
                    "fp32": {"achieved_tflops": dom["tflops_fp32"], "peak_tflops": round(fp32_peak, 1),
                             "frac": round(dom["tflops_fp32"] / fp32_peak, 4),
                             "note": "MR-STFT loss is fp32-pipe bound (~75 flop/B against an fp32 ridge of 11.4 flop/B, SURVEY 8d): "
                                     "the HBM fraction of this kernel cannot be high; flops = 2.5 n log2 n per transform actually run"}}
        for k in kernels:
            k["hbm_frac"] = round(k["gbs"] / hbm_peak, 4)

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        crow = min(args.cpu_rows, rows)
        dt, threads = cpu_chain(crow, N, 3, 1)
        cpu_baseline = {"value": crow * N / SR / dt, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{crow} of {rows} rows per step, 3 timed steps ({dt * 1e3:.0f} ms/step), torch CPU oracle",
                        "host": host_info()}
        if not args.no_configs:
            # the oracle as the checker (and the timed CPU arm) of every other config, bounded samples
            cpu_baseline["configs"] = cpu_configs_parity(se, dev)

    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "rounds_ms_per_step": [round(v, 5) for v in rounds_ms], "timing": "median of rounds",
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world, exchange),
        "clocks": clocks, "e2e": e2e, "configs": configs, "incumbent": incumbent, "sharded_loss_check": shard_check, "api_overhead": api_overhead, "gpu_launches": n_launch * args.steps, "launches_per_step": n_launch,
        "composition": comp, "alt_compositions": alts,
        "loss": loss_val, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline,
    }))
    if group is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
