"""The on-wire format INTO the path: `collate_fn_pad` (src/distrib.py:38-98) cuts
variable-length utterances into fixed segments on the host, and the Solver moves the batch to the
device synchronously (src/solver.py:431-432).  Here the collate writes straight into pinned staging
buffers and a copy stream uploads batch i+1 while the kernels work on batch i.

Plumbing only (PyTorch pinned memory, streams, events): no arithmetic happens here.
"""
from __future__ import annotations

import torch


def _segments_of(length, segment_length, drop_last):
    """How collate_fn_pad sizes one utterance (distrib.py:56-66): short clips are padded up to one
    segment; the remainder is dropped or zero-padded."""
    length = max(length, segment_length)
    if length % segment_length == 0:
        return length // segment_length
    return length // segment_length if drop_last else length // segment_length + 1


def collate_pad(batch, segment_length, drop_last=True, out=None):
    """batch: list of (mixture [C,L], sources [S,C,L], ...).  Returns (mixture [sum nseg, C, seg],
    sources [sum nseg, S, C, seg], index_batch) like the reference (distrib.py:84-96), written into
    `out` = (mix_buf, src_buf) when given (pinned staging) instead of freshly allocated tensors."""
    nsegs = [_segments_of(item[0].shape[-1], segment_length, drop_last) for item in batch]
    total = sum(nsegs)
    nch = batch[0][0].shape[0]
    nspk = batch[0][1].shape[0]
    if out is None:
        mix = torch.zeros(total, nch, segment_length, dtype=batch[0][0].dtype)
        src = torch.zeros(total, nspk, nch, segment_length, dtype=batch[0][1].dtype)
    else:
        if total > out[0].shape[0] or total > out[1].shape[0]:
            raise ValueError(f"batch has {total} segments but the staging buffers hold {min(out[0].shape[0], out[1].shape[0])} "
                             "(raise max_segments)")
        mix, src = out[0][:total], out[1][:total]
        mix.zero_()
        src.zero_()
    at = 0
    for item, ns in zip(batch, nsegs):
        mixture, sources = item[0], item[1]
        keep = min(mixture.shape[-1], ns * segment_length)
        full, rem = divmod(keep, segment_length)
        if full:
            mix[at:at + full] = mixture[:, :full * segment_length].reshape(nch, full, segment_length).permute(1, 0, 2)
            src[at:at + full] = sources[:, :, :full * segment_length].reshape(nspk, nch, full, segment_length).permute(2, 0, 1, 3)
        if rem:                                     # zero-padded tail segment (short clip or drop_last=False)
            mix[at + full, :, :rem] = mixture[:, full * segment_length:keep]
            src[at + full, :, :, :rem] = sources[:, :, full * segment_length:keep]
        at += ns
    return mix, src, nsegs


class PinnedFeeder:
    """Double-buffered pinned staging + copy stream.

        feeder = PinnedFeeder(loader, segment_length, device, max_segments=64, channels=1, speakers=1)
        for mixture, sources, index_batch in feeder:      # device tensors, upload of the next batch in flight
            ...

    `loader` yields lists of (mixture [C,L], sources [S,C,L], ...) items (a DataLoader with
    `collate_fn=lambda b: b`).  The yielded tensors are views of the device staging slots.  Lifetime: a batch is valid
    until the NEXT batch is requested, for work queued on the current stream up to that point -- requesting batch i+1
    records the consumer's progress and immediately queues the upload of batch i+2 into batch i's slot behind it.  Work
    on batch i that is queued later, or on another stream without an event dependency, races with that upload: clone
    what must live longer.  A batch with more than `max_segments` segments raises ValueError.
    """

    def __init__(self, loader, segment_length, device, max_segments, channels=1, speakers=1, drop_last=True):
        self.loader, self.seg, self.device, self.drop_last = loader, int(segment_length), torch.device(device), drop_last
        self.host = [(torch.empty(max_segments, channels, self.seg).pin_memory(),
                      torch.empty(max_segments, speakers, channels, self.seg).pin_memory()) for _ in range(2)]
        self.dev = [(torch.empty(max_segments, channels, self.seg, device=self.device),
                     torch.empty(max_segments, speakers, channels, self.seg, device=self.device)) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
        self.bytes_per_batch = 0

    def _upload(self, slot, batch):
        self.h2d_done[slot].synchronize()             # the pinned slot may still be read by an older copy
        mix, src, nsegs = collate_pad(batch, self.seg, self.drop_last, out=self.host[slot])
        n = mix.shape[0]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])
            self.dev[slot][0][:n].copy_(mix, non_blocking=True)
            self.dev[slot][1][:n].copy_(src, non_blocking=True)
            self.h2d_done[slot].record(self.copy_stream)
            self.ready[slot].record(self.copy_stream)
        self.bytes_per_batch = (mix.numel() + src.numel()) * 4
        return n, nsegs

    def __iter__(self):
        main = torch.cuda.current_stream(self.device)
        for s in range(2):
            self.consumed[s].record(main)
            self.h2d_done[s].record(main)
        it = iter(self.loader)
        pending = None
        try:
            pending = (0,) + self._upload(0, next(it))
        except StopIteration:
            return
        i = 0
        while pending is not None:
            slot, n, nsegs = pending
            try:
                nxt = next(it)
                pending = ((i + 1) & 1,) + self._upload((i + 1) & 1, nxt)
            except StopIteration:
                pending = None
            main = torch.cuda.current_stream(self.device)
            main.wait_event(self.ready[slot])
            yield self.dev[slot][0][:n], self.dev[slot][1][:n], nsegs
            self.consumed[slot].record(torch.cuda.current_stream(self.device))
            i += 1
