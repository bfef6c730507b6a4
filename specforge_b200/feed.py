"""Double-buffered host->device input feed (SURVEY §8f "input feed" row; the reference does a blocking pageable
`.to(device)` per micro-batch inside forward_loss, strategies/base.py:282-289).

`DevicePrefetcher` wraps any iterable of TrainBatch-like objects whose tensors live in (ideally pinned) host memory and
yields the same objects with device tensors.  The copy of batch k+1 is issued on a side stream while step k computes;
an event hands each batch over to the compute stream, so the step never waits on PCIe unless the feed is behind."""
from __future__ import annotations

import dataclasses
from typing import Iterable, Iterator, Optional

import torch


def mark_copied(host_tensor: torch.Tensor, event: "torch.cuda.Event") -> None:
    """Called by whoever enqueues a non-blocking H2D copy FROM a reusable pinned host tensor: the producer that owns the buffer must
    not overwrite it before `event` (recorded after the copy on the copying stream) has completed."""
    host_tensor._sf_copy_event = event


def wait_until_copied(host_tensor: torch.Tensor) -> None:
    """Producer side: block (host) until the last asynchronous copy out of this pinned buffer has executed."""
    ev = getattr(host_tensor, "_sf_copy_event", None)
    if ev is not None:
        ev.synchronize()
        host_tensor._sf_copy_event = None


_FEED_STREAMS = {}


def feed_stream(device: torch.device) -> "torch.cuda.Stream":
    """ONE copy stream per device for every prefetcher of the process.  The caching allocator keeps a pool per stream: a prefetcher
    that made its own stream (one per epoch, say) had to cudaMalloc its staging tensors again — half a gigabyte per batch at the
    BASELINE shapes — and the freed blocks of the previous stream's pool could not serve it."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _FEED_STREAMS.get(idx)
    if st is None:
        st = _FEED_STREAMS[idx] = torch.cuda.Stream(device=device)
    return st


class DevicePrefetcher:
    def __init__(self, batches: Iterable, device: Optional[torch.device] = None, depth: int = 2):
        self.batches = batches
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.depth = max(1, depth)
        self.stream = feed_stream(self.device)

    def _stage(self, batch):
        with torch.cuda.stream(self.stream):
            tensors, sources = {}, []
            for k, v in batch.tensors.items():
                if isinstance(v, torch.Tensor) and not v.is_cuda:
                    if not v.is_pinned():
                        v = v.pin_memory()
                    else:
                        sources.append(v)
                    tensors[k] = v.to(self.device, non_blocking=True)
                else:
                    tensors[k] = v
            ev = torch.cuda.Event()
            ev.record(self.stream)
            for v in sources:                 # a loader that reuses its pinned buffers waits on this before rewriting them
                mark_copied(v, ev)
        return dataclasses.replace(batch, tensors=tensors), ev

    def __iter__(self) -> Iterator:
        it = iter(self.batches)
        queue = []
        try:
            for _ in range(self.depth):
                queue.append(self._stage(next(it)))
        except StopIteration:
            pass
        while queue:
            batch, ev = queue.pop(0)
            torch.cuda.current_stream(self.device).wait_event(ev)
            for v in batch.tensors.values():          # the compute stream now owns these buffers
                if isinstance(v, torch.Tensor) and v.is_cuda:
                    v.record_stream(torch.cuda.current_stream(self.device))
            try:
                queue.append(self._stage(next(it)))
            except StopIteration:
                pass
            yield batch
