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
    """In-place SUM all-reduce of the [9] float64 partial sums (NCCL on GPU, gloo in CPU tests)."""
    import torch.distributed as dist
    if sums.dtype != torch.float64 or sums.numel() != 9:
        raise ValueError("expected the 9 float64 MR-STFT partial sums")
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def loss_from_sums(sums, global_rows: int, nsample: int) -> float:
    """Host restatement of se_mrstft_loss_value (csrc/se_kernels.cuh k_loss_value)."""
    total = 0.0
    for r, (n_fft, hop, _) in enumerate(RESOLUTIONS):
        d2, b2, lm = (float(v) for v in sums[3 * r: 3 * r + 3])
        count = global_rows * (n_fft // 2 + 1) * (1 + nsample // hop)
        total += math.sqrt(d2) / math.sqrt(b2) + lm / count
    return total / len(RESOLUTIONS)
