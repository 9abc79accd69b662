"""Host -> device prefetch of training batches on a side stream.

The reference's loop (``src/training/train.py:226-247``) moves each batch to the device and then runs the step; with
a 49 MB log-prob batch the copy (PCIe) takes longer than the fused loss launch, so the two are overlapped: while
step i runs, step i+1's tensors are already on their way.  This is the usual pinned-memory double buffer; nothing
here touches the arithmetic.
"""

from __future__ import annotations

from typing import Iterable, Iterator, Optional, Sequence, Tuple

import torch


class DevicePrefetcher:
    """Iterates ``source`` (tuples of host tensors, ideally pinned) one batch ahead on a copy stream.

    ``next()`` returns batch i as device tensors that are safe to use on the CURRENT stream, after having queued
    the copies of batch i+1."""

    def __init__(self, source: Iterable[Sequence[torch.Tensor]], device: torch.device):
        self.it: Iterator[Sequence[torch.Tensor]] = iter(source)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._next: Optional[Tuple[Tuple[torch.Tensor, ...], torch.cuda.Event]] = None
        self._queue()

    def _queue(self) -> None:
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.copy_stream):
            dev = tuple(t.to(self.device, non_blocking=True) if isinstance(t, torch.Tensor) else t for t in host)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._next = (dev, ev)

    def __iter__(self):
        return self

    def __next__(self) -> Tuple[torch.Tensor, ...]:
        if self._next is None:
            raise StopIteration
        dev, ev = self._next
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev:
            if isinstance(t, torch.Tensor):
                t.record_stream(cur)  # allocated on the copy stream, consumed on this one
        self._queue()
        return dev
