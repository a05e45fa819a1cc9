"""Per-kernel CUDA-event timing of the rasterizer stages (include/texgs.h TEXGS_EV_*).

Usage:
    timer = StageTimer(capacity=64)
    with timer.view():            # one forward (+ its backward) records into one slot
        pkg = uv_tex_render(...); loss.backward()
    torch.cuda.synchronize(); timer.summary() -> {stage: mean ms}
Events are recorded on the stream the kernels are launched on, by the library itself, so the
durations are device times of exactly those launches.
"""
from __future__ import annotations

import ctypes as C
import threading
from contextlib import contextmanager

import torch

from . import _lib as L

_tls = threading.local()


def current_event_array():
    """ctypes array (TEXGS_EV_COUNT void*) for the calling thread's active slot, or None."""
    return getattr(_tls, "arr", None)


class StageTimer:
    STAGES = (("preprocess_fwd", 0, 1), ("scan_tiles", 1, 2), ("scatter_pairs", 2, 3), ("sort_tiles", 3, 4),
              ("render_fwd", 4, 5), ("bwd_clear", 6, 7), ("render_bwd", 7, 8), ("preprocess_bwd", 8, 9),
              ("forward_total", 0, 5), ("backward_total", 6, 9))

    def __init__(self, capacity: int = 64, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.slots = []
        with torch.cuda.device(self.device):
            for _ in range(capacity):
                evs = [torch.cuda.Event(enable_timing=True) for _ in range(L.EV_COUNT)]
                for e in evs:
                    e.record()            # forces creation of the cudaEvent_t handle
                arr = (C.c_void_p * L.EV_COUNT)(*[C.c_void_p(e.cuda_event) for e in evs])
                self.slots.append((evs, arr))
            torch.cuda.synchronize(self.device)
        self.used = 0
        self.has_bwd = []

    @contextmanager
    def view(self, backward: bool = True):
        if self.used >= len(self.slots):
            yield            # out of slots: run untimed
            return
        evs, arr = self.slots[self.used]
        self.used += 1
        self.has_bwd.append(backward)
        _tls.arr = arr
        try:
            yield
        finally:
            _tls.arr = None

    def summary(self) -> dict:
        """Mean milliseconds per stage over the recorded views (call after a synchronize)."""
        out = {}
        for name, a, b in self.STAGES:
            vals = []
            for i in range(self.used):
                if a >= 6 and not self.has_bwd[i]:
                    continue
                evs = self.slots[i][0]
                vals.append(evs[a].elapsed_time(evs[b]))
            if vals:
                out[name] = sum(vals) / len(vals)
        return out

    def reset(self):
        self.used = 0
        self.has_bwd = []
