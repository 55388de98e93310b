"""Double-buffered host -> device staging of per-view inputs (camera matrices, target image) on a copy stream.

The rasteriser step of one view takes a few milliseconds; a 1080p float target image is 25 MB, i.e. about a
millisecond of PCIe time.  Issued on the compute stream the copy serialises with the kernels; issued one view ahead
on its own stream it disappears underneath them.  The reference trainer keeps its single target image resident
(example.py:62-66) and has no multi-view loader; this is the loader a multi-view caller of the drop-in API needs.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


class ViewPrefetcher:
    """Ring of ``depth`` device slots filled from pinned host tensors on a dedicated copy stream.

    ``submit(host_tensors)`` enqueues the copies of the NEXT view; ``get()`` returns the device tensors of the oldest
    submitted view (the current stream waits for its copy); ``release()`` tells the ring that the current stream has
    finished reading that view, so its slot may be overwritten by a later ``submit``."""

    def __init__(self, device: torch.device, depth: int = 2):
        if torch.device(device).type != "cuda":
            raise RuntimeError("ViewPrefetcher stages into CUDA memory (there is no CPU path)")
        self.dev = torch.device(device)
        self.depth = int(depth)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.slots: List[Tuple[torch.Tensor, ...]] = [() for _ in range(self.depth)]
        self.filled = [torch.cuda.Event() for _ in range(self.depth)]
        self.free = [torch.cuda.Event() for _ in range(self.depth)]
        self._free_valid = [False] * self.depth
        self.head = 0   # next slot to fill
        self.tail = 0   # next slot to hand out
        self.pending = 0
        self.bytes_copied = 0

    def submit(self, host_tensors: Sequence[torch.Tensor]) -> None:
        if self.pending >= self.depth:
            raise RuntimeError("ViewPrefetcher ring is full: get()/release() a view before submitting another")
        k = self.head
        for t in host_tensors:
            if not t.is_pinned():
                raise RuntimeError("ViewPrefetcher expects pinned host tensors (pin_memory())")
        if not self.slots[k] or any(a.shape != b.shape or a.dtype != b.dtype for a, b in zip(self.slots[k], host_tensors)):
            self.slots[k] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.dev) for t in host_tensors)
        with torch.cuda.stream(self.copy_stream):
            if self._free_valid[k]:
                self.copy_stream.wait_event(self.free[k])  # the last reader of this slot is done
            for d, h in zip(self.slots[k], host_tensors):
                d.copy_(h, non_blocking=True)
                self.bytes_copied += h.numel() * h.element_size()
            self.filled[k].record(self.copy_stream)
        self.head = (k + 1) % self.depth
        self.pending += 1

    def get(self) -> Tuple[torch.Tensor, ...]:
        if self.pending == 0:
            raise RuntimeError("ViewPrefetcher.get() without a submitted view")
        k = self.tail
        torch.cuda.current_stream(self.dev).wait_event(self.filled[k])
        return self.slots[k]

    def release(self) -> None:
        k = self.tail
        self.free[k].record(torch.cuda.current_stream(self.dev))
        self._free_valid[k] = True
        self.tail = (k + 1) % self.depth
        self.pending -= 1
