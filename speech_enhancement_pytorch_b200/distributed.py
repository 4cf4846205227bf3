"""Utterance-sharded multi-GPU plumbing.

Rows of the path are independent, so ranks take contiguous blocks of utterances and run the
kernels with no data-path collective.  The MR-STFT loss needs ONE exchange step: the 9 partial
sums (per resolution: sum (b-a)^2, sum b^2, sum |log b - log a|) are all-reduced so that the
batch-global Frobenius ratio and mean equal the single-GPU result.  The reference's multi-GPU
path is nn.DataParallel (src/solver.py:144-145) and never parallelises the STFT at all.
"""
from __future__ import annotations

import math

import torch

RESOLUTIONS = ((512, 128, 512), (1024, 256, 1024), (2048, 512, 2048))


def shard_rows(n_utterances: int, world: int, rank: int) -> slice:
    """Contiguous block of utterances for `rank` (all channels/speakers of an utterance stay together)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(n_utterances, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def all_reduce_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of the [9] float64 partial sums, or of [10] = the sums + this rank's row count (the
    exchange then leaves the global row count in [9]: uneven shards).  NCCL on GPU, gloo in CPU tests."""
    import torch.distributed as dist
    if sums.dtype != torch.float64 or sums.numel() not in (9, 10):
        raise ValueError("expected the 9 (or 9 + row count) float64 MR-STFT partial sums")
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


class PeerExchange:
    """This rank's view of the peer-memory exchange buffers of a process group (csrc/se_api_p2p.cu).

    `exchange_value` is the fused replacement of `all_reduce_sums` + `se_mrstft_loss_value`: one single-CTA kernel on
    the current stream stores the 9 sums into every peer's buffer over NVLink, waits on the device for all ranks,
    adds in rank order and writes the global sums (in place) and the loss."""

    def __init__(self, group, device):
        import ctypes
        import socket
        import torch.distributed as dist
        from . import _native as nv
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise NotImplementedError("peer exchange supports up to 16 ranks")
        self.device = torch.device(device)
        L = nv.lib()
        local = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        self.local, self.peers, err = None, [], None
        try:
            with nv.on_device(self.device):
                nv.check(L.se_p2p_create(ctypes.byref(local), handle))
            self.local = local.value
        except Exception as e:                     # still take part in the gather below, then fail with everyone
            err = e
        infos = [None] * self.world
        dist.all_gather_object(infos, (socket.gethostname(), handle.raw if err is None else None), group=group)
        ptrs = (ctypes.c_void_p * self.world)()
        try:
            if err is not None:
                raise err
            if any(raw is None for _, raw in infos):
                raise NotImplementedError("a rank of the group could not create its exchange buffer")
            if len({h for h, _ in infos}) != 1:
                raise NotImplementedError("peer exchange needs all ranks of the group on one node")
            for r, (_, raw) in enumerate(infos):
                if r == self.rank:
                    ptrs[r] = self.local
                    continue
                peer = ctypes.c_void_p()
                with nv.on_device(self.device):
                    nv.check(L.se_p2p_open(raw, ctypes.byref(peer)))
                self.peers.append(peer.value)
                ptrs[r] = peer.value
        except Exception:
            self.close()
            raise
        self.ptrs = ptrs

    def exchange_value(self, sums: torch.Tensor, global_rows, nsample: int, loss, stream: int):
        """sums [9] with a host-known global_rows, or sums [10] (global_rows None): [9] carries this rank's row count
        in and the global count out, so uneven shards need no guess."""
        from . import _native as nv
        if sums.dtype != torch.float64 or sums.numel() not in (9, 10) or sums.device != self.device:
            raise ValueError("expected the 9 (or 9 + row count) float64 MR-STFT partial sums on this exchange's device")
        lp = loss.data_ptr() if loss is not None else None
        if sums.numel() == 10:
            nv.check(nv.lib().se_mrstft_exchange_rows_value(sums.data_ptr(), self.ptrs, self.world, self.rank, nsample, lp, stream))
        else:
            nv.check(nv.lib().se_mrstft_exchange_value(sums.data_ptr(), self.ptrs, self.world, self.rank, global_rows, nsample, lp, stream))

    def close(self):
        from . import _native as nv
        L = nv.lib()
        with nv.on_device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self.peers:
                L.se_p2p_close(p)
            self.peers = []
            if self.local:
                L.se_p2p_destroy(self.local)
                self.local = None


# One PeerExchange per (process group, device).  Keyed on the group OBJECT through a WeakKeyDictionary: destroying a
# group drops its entry (and closes the IPC mappings) instead of leaving a stale one behind an id() that a later group
# may reuse.
import weakref

_exchanges = weakref.WeakKeyDictionary()


def close_exchanges(group=None):
    """Close the peer exchanges of `group` (or of every group): call before dist.destroy_process_group()."""
    groups = [group] if group is not None else list(_exchanges.keys())
    for g in groups:
        for px in (_exchanges.pop(g, None) or {}).values():
            if px is not None:
                px.close()


def peer_exchange(group, device):
    """The group's PeerExchange on `device`, created on first use (a collective call: every rank of the group must
    reach it).  Returns None -- and the caller uses NCCL -- when SE_P2P_EXCHANGE=0, when the ranks are not all on one
    node, or when any rank could not map its peers; the decision is taken jointly so that all ranks agree."""
    import os
    import torch.distributed as dist
    per_group = _exchanges.setdefault(group, {})
    key = torch.device(device).index
    if key in per_group:
        return per_group[key]
    px, ok = None, os.environ.get("SE_P2P_EXCHANGE", "1") != "0"
    if ok:
        try:
            px = PeerExchange(group, device)
        except NotImplementedError:
            ok = False
        except Exception as e:                     # e.g. no peer access between two of the GPUs
            import warnings
            warnings.warn(f"peer exchange unavailable on rank {dist.get_rank(group)} ({e}); using NCCL")
            ok = False
    flags = [None] * dist.get_world_size(group)
    dist.all_gather_object(flags, ok, group=group)
    if not all(flags):
        if px is not None:
            px.close()
        px = None
    per_group[key] = px
    return px


def loss_from_sums(sums, global_rows: int, nsample: int) -> float:
    """Host restatement of se_mrstft_loss_value (csrc/se_kernels.cuh k_loss_value)."""
    total = 0.0
    for r, (n_fft, hop, _) in enumerate(RESOLUTIONS):
        d2, b2, lm = (float(v) for v in sums[3 * r: 3 * r + 3])
        count = global_rows * (n_fft // 2 + 1) * (1 + nsample // hop)
        total += math.sqrt(d2) / math.sqrt(b2) + lm / count
    return total / len(RESOLUTIONS)
