"""Fold lanes: independent units of a driver (the folds of a ``Validator``, the (set, fold) units of ``Hierarchical``)
enqueued on a few CUDA streams from one host thread each.

The reference hands such units to joblib workers (validator.py:365-380, hierarchical.py:468).  On one GPU the reason to
overlap them is different: a large fit is a chain of ~5 000 dependent launches, many of which cannot fill 148 SMs (the
block-Lanczos steps run 1-CTA factorisations and 0.9 / 1.7-wave skinny products, the block Cholesky runs 20 CTAs), so
a second unit's kernels fill the SMs the first leaves idle.  Same kernels, same per-unit launch order, same numbers.

Discipline that makes the caching allocator and cross-stream reads safe without ``record_stream``:
  * every lane waits for the caller's stream before its first task, and the caller's stream waits for every lane after
    the last one, so nothing crosses between lanes and the caller inside a call;
  * tasks with the same ``lane_key`` always run on the same lane, in order (shared per-fold statistics stay on one stream).
One host thread per lane (a small persistent pool), because the pending-launch queue of a stream is bounded: a single thread
enqueueing a whole unit on lane 0 would block before it reaches lane 1.  ctypes calls and torch ops release the GIL.
"""
from __future__ import annotations

import os
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Callable, Dict, List, Optional, Sequence

import torch

_lanes: Dict[int, List['torch.cuda.Stream']] = {}
_n_override: Optional[int] = None
_lock = threading.Lock()
MAX_LANES = 8
_pool: Optional[ThreadPoolExecutor] = None      # persistent lane threads: the library keeps a few per-thread CUDA events, so threads are reused
_in_lane = threading.local()


def _lane_pool() -> ThreadPoolExecutor:
    global _pool
    with _lock:
        if _pool is None:
            _pool = ThreadPoolExecutor(max_workers=MAX_LANES, thread_name_prefix='mvb-lane')
        return _pool


def set_fold_streams(n: Optional[int]) -> None:
    """Number of lanes (1 = serial on the caller's stream); None restores the default / ``MVB_FOLD_STREAMS``."""
    global _n_override
    _n_override = None if n is None else min(MAX_LANES, max(1, int(n)))


def n_fold_streams() -> int:
    if _n_override is not None:
        return _n_override
    return min(MAX_LANES, max(1, int(os.environ.get('MVB_FOLD_STREAMS', '3'))))


def _device_lanes(device: torch.device, n: int) -> List['torch.cuda.Stream']:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    with _lock:
        lanes = _lanes.setdefault(idx, [])
        while len(lanes) < n:
            lanes.append(torch.cuda.Stream(device=idx))
        return lanes[:n]


_warmed = False


def _warm(device: torch.device) -> None:
    """torch loads its linalg backend lazily and that first call is not thread-safe ("lazy wrapper should be called at
    most once"): make it here, once, from the caller's thread (TimeDelayed's patterns invert an f x f covariance)."""
    global _warmed
    if not _warmed:
        torch.linalg.inv(torch.eye(2, device=device))
        _warmed = True


def run_on_lanes(tasks: Sequence[Callable[[], Any]], device: torch.device, lane_keys: Optional[Sequence[int]] = None) -> List[Any]:
    """Run ``tasks`` (zero-argument callables that enqueue CUDA work on the *current* stream) and return their results in
    order.  ``lane_keys[i]`` (default i) picks the lane: ``key % n_lanes``."""
    from . import _lib
    n = min(n_fold_streams(), len(tasks))
    if n <= 1 or device.type != 'cuda' or getattr(_in_lane, 'active', False) or torch.cuda.is_current_stream_capturing():
        return [t() for t in tasks]        # (a driver nested inside a lane task stays on that lane: the pool is not re-entered)
    keys = list(range(len(tasks))) if lane_keys is None else [int(k) for k in lane_keys]
    _warm(device)
    lanes = _device_lanes(device, n)
    main = torch.cuda.current_stream(device)
    assigned: List[List[int]] = [[] for _ in range(n)]
    for i, k in enumerate(keys):
        assigned[k % n].append(i)
    results: List[Any] = [None] * len(tasks)
    errors: List[Optional[BaseException]] = [None] * n
    parent = _lib.fold_ctx.parent

    def work(lane: int) -> None:
        try:
            _in_lane.active = True
            torch.cuda.set_device(device)
            _lib.fold_ctx.parent = parent
            with torch.cuda.stream(lanes[lane]):
                for i in assigned[lane]:
                    results[i] = tasks[i]()
        except BaseException as e:          # re-raised in the caller's thread
            errors[lane] = e
        finally:
            _lib.fold_ctx.parent = None
            _in_lane.active = False

    for s in lanes:
        s.wait_stream(main)
    futures = [_lane_pool().submit(work, lane) for lane in range(n) if assigned[lane]]
    for f in futures:
        f.result()
    for s in lanes:
        main.wait_stream(s)
    for e in errors:
        if e is not None:
            raise e
    return results
