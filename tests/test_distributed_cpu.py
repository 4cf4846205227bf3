"""N>1 host logic on CPU: world_size-2 gloo.  Each rank takes its utterance shard, forms the 9
partial sums (with the oracle standing in for the kernels -- there is no GPU here), all-reduces
them, and must reproduce the single-process batch-global loss."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from speech_enhancement_pytorch_b200 import distributed as sed


def test_shard_rows_partitions_exactly():
    for n in (1, 7, 64, 129):
        for world in (1, 2, 3, 8):
            parts = [sed.shard_rows(n, world, r) for r in range(world)]
            covered = [i for p in parts for i in range(p.start, p.stop)]
            assert covered == list(range(n))
            sizes = [p.stop - p.start for p in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sed.shard_rows(4, 2, 2)


def _worker(rank, world, port, nsample, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import spectral_oracle as oref
    g = torch.Generator().manual_seed(4)
    ref = torch.randn(7, 1, nsample, generator=g)
    est = ref + 0.2 * torch.randn(7, 1, nsample, generator=g)
    sl = sed.shard_rows(7, world, rank)                  # UNEVEN shards: 4 + 3 rows
    parts = oref.mrstft_partials_ref(est[sl], ref[sl])
    # the exchange carries the 9 sums AND this rank's row count: no rank guesses the global count as rows * world
    sums = torch.tensor([v for p in parts for v in p[:3]] + [float(sl.stop - sl.start)], dtype=torch.float64)
    sed.all_reduce_sums(sums)
    assert int(sums[9]) == 7
    out[rank] = sed.loss_from_sums(sums[:9], int(sums[9]), nsample)
    dist.destroy_process_group()


def test_two_rank_loss_equals_global_loss():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    nsample = 5000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, nsample, out), nprocs=2, join=True)
    from oracle import spectral_oracle as oref
    g = torch.Generator().manual_seed(4)
    ref = torch.randn(7, 1, nsample, generator=g)
    est = ref + 0.2 * torch.randn(7, 1, nsample, generator=g)
    want = float(oref.mrstft_loss_ref(est, ref))
    assert abs(out[0] - out[1]) < 1e-12
    assert abs(out[0] - want) / want < 1e-5


def _exchange_worker(rank, world, port, disabled, out):
    import warnings
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SE_P2P_EXCHANGE="0" if disabled else "1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    with warnings.catch_warnings(record=True) as seen:
        warnings.simplefilter("always")
        px = sed.peer_exchange(dist.group.WORLD, torch.device("cuda", 0))
        again = sed.peer_exchange(dist.group.WORLD, torch.device("cuda", 0))      # cached decision, no second gather
    out[rank] = (px is None, again is None, [str(w.message) for w in seen])
    dist.destroy_process_group()


@pytest.mark.parametrize("disabled", [False, True])
def test_peer_exchange_falls_back_jointly_without_peer_memory(disabled):
    """No GPU here, so no rank can create its exchange buffer: every rank must still take part in the two gathers,
    agree on `None` (the caller then uses the NCCL / gloo all-reduce) and say why -- no hang, no exception.  With
    SE_P2P_EXCHANGE=0 the buffers are not even attempted."""
    if torch.cuda.is_available():
        pytest.skip("the fallback decision is exercised on machines without a GPU")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_exchange_worker, args=(2, port, disabled, out), nprocs=2, join=True)
    for r in range(2):
        none1, none2, msgs = out[r]
        assert none1 and none2
        assert (msgs == []) if disabled else any("peer exchange unavailable" in m for m in msgs)
