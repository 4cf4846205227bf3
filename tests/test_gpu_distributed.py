"""Two-rank check of the path's only exchange step (needs >= 2 GPUs; skipped otherwise): utterance-sharded
loss_mrstft(group=WORLD) must equal the single-GPU batch-global loss and gradient -- through the fused peer-memory
exchange kernel (default) and through NCCL (SE_P2P_EXCHANGE=0) -- with UNEVEN shards (7 rows over 2 ranks: the row
counts travel with the sums, nothing is guessed on the host).  The STFT-domain mse loss is checked the same way."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out, p2p):
    import torch.distributed as dist
    import speech_enhancement_pytorch_b200 as se
    from speech_enhancement_pytorch_b200.distributed import shard_rows, peer_exchange
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NCCL_DEBUG="WARN", SE_P2P_EXCHANGE="1" if p2p else "0")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    g = torch.Generator().manual_seed(8)
    ref = torch.randn(7, 1, 16000, generator=g)
    est = ref + 0.2 * torch.randn(7, 1, 16000, generator=g)
    sl = shard_rows(7, world, rank)                       # 4 + 3 rows
    e = est[sl].cuda().requires_grad_(True)
    loss = se.loss_mrstft(e, ref[sl].cuda(), group=dist.group.WORLD)
    loss.backward()
    used = peer_exchange(dist.group.WORLD, e.device) is not None
    # the exchange buffers alternate between two parities: repeated steps must keep giving the same bits
    again = [float(se.loss_mrstft(e.detach(), ref[sl].cuda(), group=dist.group.WORLD)) for _ in range(5)]
    # DDP averages gradients over ranks: scale_grad_by_world compensates (x world), the loss value is unchanged
    e2 = est[sl].cuda().requires_grad_(True)
    l2 = se.loss_mrstft(e2, ref[sl].cuda(), group=dist.group.WORLD, scale_grad_by_world=True)
    l2.backward()
    scaled_ok = float(l2) == float(loss) and float((e2.grad - world * e.grad).abs().max()) <= 1e-6 * float(e.grad.abs().max())
    # STFT-domain mse against a waveform target, uneven shards
    import types
    cfg = types.SimpleNamespace(n_fft=512, hop_length=128, win_length=512, center=True)
    spec = se.stft_custom(est[sl].cuda(), cfg).detach().requires_grad_(True)
    ls = se.loss_spectral(spec, ref[sl].cuda(), cfg, "mse", group=dist.group.WORLD)
    ls.backward()
    out[rank] = (float(loss), e.grad.cpu(), used, again, scaled_ok, float(ls), spec.grad.cpu())
    from speech_enhancement_pytorch_b200.distributed import close_exchanges
    close_exchanges(dist.group.WORLD)
    dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [True, False])
def test_two_rank_loss_matches_single_gpu(p2p):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import speech_enhancement_pytorch_b200 as se
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out, p2p), nprocs=2, join=True)
    g = torch.Generator().manual_seed(8)
    ref = torch.randn(7, 1, 16000, generator=g)
    est = ref + 0.2 * torch.randn(7, 1, 16000, generator=g)
    e = est.cuda().requires_grad_(True)
    loss = se.loss_mrstft(e, ref.cuda())
    loss.backward()
    assert out[0][2] == p2p and out[1][2] == p2p          # the peer-memory kernel really ran (or really did not)
    assert out[0][0] == out[1][0]                         # rank-ordered sum: identical bits on every rank
    assert all(v == out[0][0] for v in out[0][3] + out[1][3])
    assert abs(out[0][0] - float(loss)) < 1e-6 * abs(float(loss)) + 1e-7
    assert abs(out[1][0] - float(loss)) < 1e-6 * abs(float(loss)) + 1e-7
    grad = torch.cat([out[0][1], out[1][1]], 0)
    assert float((grad - e.grad.cpu()).abs().max()) < 1e-6 * float(e.grad.abs().max()) + 1e-9
    assert out[0][4] and out[1][4]
    import types
    cfg = types.SimpleNamespace(n_fft=512, hop_length=128, win_length=512, center=True)
    spec = se.stft_custom(est.cuda(), cfg).detach().requires_grad_(True)
    ls = se.loss_spectral(spec, ref.cuda(), cfg, "mse")
    ls.backward()
    assert abs(out[0][5] - float(ls)) < 1e-6 * abs(float(ls)) and abs(out[1][5] - float(ls)) < 1e-6 * abs(float(ls))
    gs = torch.cat([out[0][6], out[1][6]], 0)
    assert float((gs - spec.grad.cpu()).abs().max()) < 1e-6 * float(spec.grad.abs().max()) + 1e-12
