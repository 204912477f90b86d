"""CUDA-graph execution of many small, identically shaped fits (``Sliding`` over a ridge estimator).

A per-timepoint decoder fit at BASELINE configs[4] size (1600 x 306 -> 3 targets) is ~180 dependent launches of tiny
kernels: 2.9 ms of GPU time on a handful of SMs and 2 ms of host time per slice, 2500 times per cross-validation.
Here the whole launch sequence of one ``fit`` is stream-captured once per lane into a CUDA graph over static input
buffers; every slice is then one strided copy into the buffers, one graph replay and a snapshot of the outputs, with
the lanes replaying concurrently on their own streams so the small kernels of different slices fill the GPU together.
The kernels, their order and their arithmetic are exactly those of the eager path (parity is unchanged by
construction); lanes are cached per problem signature, so the five folds of a cross-validation share them.

Environment: ``MVB_SLIDING_GRAPHS=0`` switches this off, ``MVB_SLIDING_LANES`` sets the number of lanes (default 8)."""
from __future__ import annotations

import copy
import os
import warnings
from collections import OrderedDict
from typing import List, Optional, Sequence, Tuple

import sklearn.base
import torch

from . import _lib

_MAX_ELEMS = 1 << 22          # n * p per slice: beyond this a fit is no longer launch-bound
_MAX_COLS = 1024
_CACHE: 'OrderedDict[tuple, List[_Lane]]' = OrderedDict()
_CACHE_MAX = 4
_broken = False


def _snapshot(obj, consumer: torch.cuda.Stream):
    """Copy of a fitted estimator whose CUDA tensors no longer alias the lane's static graph outputs."""
    new = copy.copy(obj)
    for name, val in vars(obj).items():
        if isinstance(val, torch.Tensor) and val.is_cuda:
            own = val.clone()
            own.record_stream(consumer)          # produced on the lane's stream, consumed on the caller's
            setattr(new, name, own)
        elif isinstance(val, sklearn.base.BaseEstimator):
            setattr(new, name, _snapshot(val, consumer))
    return new


class _Lane:
    """One stream, one pair of static input buffers, one estimator whose ``fit`` lives in a captured graph."""

    def __init__(self, proto, x_shape: Tuple[int, ...], y_shape: Tuple[int, ...], dtype, device):
        self.stream = torch.cuda.Stream(device=device)
        self.x = torch.zeros(x_shape, dtype=dtype, device=device)
        self.y = torch.zeros(y_shape, dtype=dtype, device=device)
        self.est = proto.clone()
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches = 0

    def build(self, x0: torch.Tensor, y0: torch.Tensor) -> None:
        lib = _lib.load()
        with torch.cuda.stream(self.stream):
            self.x.copy_(x0)
            self.y.copy_(y0)
            self.est.fit(self.x, self.y)         # eager once: moves alphas to the device, sets kernel attributes, warms the allocator
        self.stream.synchronize()
        was_profiling = lib.mvb_profile_set(0)   # timing events cannot be recorded inside a capture
        try:
            n0 = _lib.launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(self.stream):         # (torch.cuda.graph() would add a gc pass and a cache flush per lane)
                self.graph.capture_begin()
                try:
                    self.est.fit(self.x, self.y)
                finally:
                    self.graph.capture_end()
            self.launches = _lib.launch_count() - n0
        finally:
            lib.mvb_profile_set(was_profiling)

    def run(self, xs: torch.Tensor, ys: torch.Tensor, consumer: torch.cuda.Stream):
        with torch.cuda.stream(self.stream):
            self.x.copy_(xs)
            self.y.copy_(ys)
            self.graph.replay()
            out = _snapshot(self.est, consumer)
        _lib.load().mvb_launch_count_add(self.launches)
        return out


def _signature(proto, xs: torch.Tensor, ys: torch.Tensor):
    from .estimators.ridgecv import _RidgeCV_torch
    from .estimators.ridgedecoder import _RidgeDecoder_torch
    if isinstance(proto, _RidgeDecoder_torch):
        alphas, opts = proto.alpha, (proto.fit_intercept, proto.alpha_per_target)
    elif isinstance(proto, _RidgeCV_torch):
        alphas, opts = proto.alphas, (proto.fit_intercept, proto.alpha_per_target)
    else:
        return None
    if not isinstance(alphas, torch.Tensor):
        return None
    return (type(proto).__name__, tuple(xs.shape), tuple(ys.shape), xs.dtype, xs.device.index, opts,
            tuple(float(a) for a in alphas.detach().cpu().reshape(-1).tolist()))


def fit_slices(proto, Xm: torch.Tensor, ym: torch.Tensor, pairs: Sequence[Tuple[int, int]]):
    """Fit ``proto.clone()`` on (Xm[..., i], ym[..., j]) for every pair; returns the list of fitted estimators, or None
    when the problem is not one this executor handles (the caller then runs its eager loop)."""
    global _broken
    if _broken or os.environ.get('MVB_SLIDING_GRAPHS', '1') == '0':
        return None
    from . import _streams
    if getattr(_streams._in_lane, 'active', False):          # stream capture next to other lanes' launches: keep the eager loop
        return None
    if not (isinstance(Xm, torch.Tensor) and Xm.is_cuda and ym.is_cuda and Xm.ndim == 3 and ym.ndim in (2, 3)):
        return None
    if Xm.dtype not in (torch.float32, torch.float64) or ym.dtype != Xm.dtype:
        return None
    n_lanes = max(1, int(os.environ.get('MVB_SLIDING_LANES', '8')))
    if len(pairs) < 2 * n_lanes or torch.cuda.is_current_stream_capturing():
        return None
    n, p = Xm.shape[0], Xm.shape[1]
    if n * p > _MAX_ELEMS or p > _MAX_COLS or n <= p + 1:
        return None
    x0, y0 = Xm[..., pairs[0][0]], ym[..., pairs[0][1]]
    key = _signature(proto, x0, y0)
    if key is None:
        return None
    caller = torch.cuda.current_stream(Xm.device)
    lanes = _CACHE.get(key)
    if lanes is None:
        try:
            lanes = [_Lane(proto, tuple(x0.shape), tuple(y0.shape), Xm.dtype, Xm.device) for _ in range(n_lanes)]
            caller.synchronize()
            for lane in lanes:
                lane.build(x0, y0)
        except Exception as e:                    # capture not possible here: keep the (identical) eager path
            _broken = True
            warnings.warn(f'mvpy_b200: CUDA-graph slice executor disabled ({type(e).__name__}: {e}); using the eager loop.')
            torch.cuda.synchronize()
            return None
        _CACHE[key] = lanes
        while len(_CACHE) > _CACHE_MAX:
            _CACHE.popitem(last=False)
    else:
        _CACHE.move_to_end(key)
    for lane in lanes:
        lane.stream.wait_stream(caller)
    fitted = [lanes[k % n_lanes].run(Xm[..., i], ym[..., j], caller) for k, (i, j) in enumerate(pairs)]
    for lane in lanes:
        caller.wait_stream(lane.stream)
    return fitted


def clear_cache() -> None:
    _CACHE.clear()
