"""Stage timing helpers -- the role of the reference's ``gsplat/profile.py:1-59`` (``timeit`` context manager /
decorator controlled by the ``TIMEIT`` environment variable, accumulating seconds into the module dict ``profiler``).

Same surface (``timeit(name)``, ``profiler``), different clock: the reference brackets a region with
``torch.cuda.synchronize()`` + ``time.perf_counter()``, which drains the launch queue twice per region and so
changes what it measures; here a region is bracketed by two CUDA events on the current stream and only read
(``elapsed_time``) when ``profiler_summary()`` / ``profiler`` is consulted, so an instrumented step runs at full
speed.  For per-kernel attribution use ``_cabi.enable_stage_timing`` (bench.py's ``stage_ms``) or set
``GSB200_NVTX=1`` to get one NVTX range per C-ABI call for Nsight.
"""
from __future__ import annotations

import os
from functools import wraps
from typing import Callable, Dict, List, Tuple

import torch


class _Profiler(dict):
    """name -> accumulated seconds; pending CUDA-event pairs are folded in on access."""

    def __init__(self):
        super().__init__()
        self._pending: List[Tuple[str, "torch.cuda.Event", "torch.cuda.Event"]] = []

    def _flush(self) -> None:
        if not self._pending:
            return
        torch.cuda.synchronize()
        for name, start, end in self._pending:
            dict.__setitem__(self, name, dict.get(self, name, 0.0) + start.elapsed_time(end) * 1e-3)
        self._pending.clear()

    def __getitem__(self, key):
        self._flush()
        return dict.__getitem__(self, key)

    def items(self):
        self._flush()
        return dict.items(self)

    def __repr__(self):
        self._flush()
        return dict.__repr__(self)

    def clear(self):
        self._pending.clear()
        dict.clear(self)


profiler: Dict[str, float] = _Profiler()


def profiler_summary() -> Dict[str, float]:
    """Accumulated seconds per region name (synchronises once to read the pending events)."""
    profiler._flush()
    return dict(profiler)


class timeit(object):
    """Times a region when the environment variable TIMEIT is 1 (otherwise a no-op).

    >>> @timeit()
    ... def my_function(): ...
    >>> with timeit(name="stage1"):
    ...     my_function()
    >>> print(profiler)
    """

    def __init__(self, name: str = "unnamed"):
        self.name = name
        self.enabled = os.environ.get("TIMEIT", "0") == "1" and torch.cuda.is_available()
        self._start = None

    def __enter__(self):
        if self.enabled:
            self._start = torch.cuda.Event(enable_timing=True)
            self._start.record()

    def __exit__(self, exc_type, exc_val, exc_tb):
        if self.enabled:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            profiler._pending.append((self.name, self._start, end))

    def __call__(self, f: Callable) -> Callable:
        @wraps(f)
        def decorated(*args, **kwargs):
            self.name = f.__name__
            with self:
                return f(*args, **kwargs)

        return decorated
