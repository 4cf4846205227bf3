"""Two-rank NCCL check of the path's only exchange step (needs >= 2 GPUs; skipped otherwise):
utterance-sharded loss_mrstft(group=WORLD) must equal the single-GPU batch-global loss and gradient."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import torch.distributed as dist
    import speech_enhancement_pytorch_b200 as se
    from speech_enhancement_pytorch_b200.distributed import shard_rows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NCCL_DEBUG="WARN")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    g = torch.Generator().manual_seed(8)
    ref = torch.randn(6, 1, 16000, generator=g)
    est = ref + 0.2 * torch.randn(6, 1, 16000, generator=g)
    sl = shard_rows(6, world, rank)
    e = est[sl].cuda().requires_grad_(True)
    loss = se.loss_mrstft(e, ref[sl].cuda(), group=dist.group.WORLD, global_rows=6)
    loss.backward()
    out[rank] = (float(loss), e.grad.cpu())
    dist.destroy_process_group()


def test_two_rank_nccl_loss_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import speech_enhancement_pytorch_b200 as se
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    g = torch.Generator().manual_seed(8)
    ref = torch.randn(6, 1, 16000, generator=g)
    est = ref + 0.2 * torch.randn(6, 1, 16000, generator=g)
    e = est.cuda().requires_grad_(True)
    loss = se.loss_mrstft(e, ref.cuda())
    loss.backward()
    assert abs(out[0][0] - float(loss)) < 1e-6 * abs(float(loss)) + 1e-7
    assert abs(out[1][0] - float(loss)) < 1e-6 * abs(float(loss)) + 1e-7
    grad = torch.cat([out[0][1], out[1][1]], 0)
    assert float((grad - e.grad.cpu()).abs().max()) < 1e-6 * float(e.grad.abs().max()) + 1e-9
