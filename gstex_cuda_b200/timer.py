"""CUDA-event phase timer, API-compatible with ``gstex_cuda/timer.py`` (start / stop / lap / dump)."""
import time

import torch


class Timer:
    def __init__(self, cuda_support=True, disabled=False):
        self.cuda_support = cuda_support and torch.cuda.is_available()
        self.disabled = disabled
        self.times = {}
        self.events = {}
        self.laps = 0
        self.key = None
        self._t0 = time.time()

    def start(self, key):
        if self.disabled:
            return
        if self.key is not None:
            self.stop()
        self.key = key
        self.times.setdefault(key, 0.0)
        if self.cuda_support:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.events.setdefault(key, []).append([ev, None])
        self._t0 = time.time()

    def stop(self):
        if self.disabled or self.key is None:
            return
        self.times[self.key] += time.time() - self._t0
        if self.cuda_support:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.events[self.key][-1][1] = ev
        self.key = None

    def lap(self):
        if not self.disabled:
            self.laps += 1

    def dump_key(self, key):
        if self.disabled:
            return
        wall_ms = 1000.0 * self.times[key]
        dev_ms = 0.0
        if self.cuda_support:
            dev_ms = sum(a.elapsed_time(b) for a, b in self.events.get(key, []) if b is not None)
            print(f"{key}\t{dev_ms:.3f}")
        else:
            print(f"{key}\t{wall_ms:.3f}")
        return wall_ms, dev_ms

    def dump(self):
        if self.disabled:
            return
        if self.cuda_support:
            torch.cuda.synchronize()
        total = 0.0
        for key in self.times:
            wall_ms, dev_ms = self.dump_key(key)
            total += dev_ms if self.cuda_support else wall_ms
        print(f"TOTAL\t{total:.3f}")
