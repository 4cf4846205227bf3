"""Two-rank check of the path's only exchange step (needs >= 2 GPUs; skipped otherwise): utterance-sharded
loss_mrstft(group=WORLD) must equal the single-GPU batch-global loss and gradient -- through the fused peer-memory
exchange kernel (default) and through NCCL (SE_P2P_EXCHANGE=0)."""
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
    ref = torch.randn(6, 1, 16000, generator=g)
    est = ref + 0.2 * torch.randn(6, 1, 16000, generator=g)
    sl = shard_rows(6, world, rank)
    e = est[sl].cuda().requires_grad_(True)
    loss = se.loss_mrstft(e, ref[sl].cuda(), group=dist.group.WORLD, global_rows=6)
    loss.backward()
    used = peer_exchange(dist.group.WORLD, e.device) is not None
    # the exchange buffers alternate between two parities: repeated steps must keep giving the same bits
    again = [float(se.loss_mrstft(e.detach(), ref[sl].cuda(), group=dist.group.WORLD, global_rows=6)) for _ in range(5)]
    out[rank] = (float(loss), e.grad.cpu(), used, again)
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
    ref = torch.randn(6, 1, 16000, generator=g)
    est = ref + 0.2 * torch.randn(6, 1, 16000, generator=g)
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
