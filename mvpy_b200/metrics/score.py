"""``metrics.score`` -- resolve what each metric requests, share one prediction, reduce, evaluate.
Same protocol as mvpy/metrics/score.py:17-122."""
from __future__ import annotations

from math import prod
from typing import Dict, Optional, Tuple, Union

import torch
from sklearn.pipeline import Pipeline

from .metric import Metric


def reduce_(X: torch.Tensor, dims: Union[int, Tuple[int, ...]]) -> torch.Tensor:
    """Move ``dims`` to the end (in ascending order) and flatten them into one trailing dim
    (mvpy/metrics/score.py:17-46).  Stays a view whenever torch allows it; the scoring kernel reads
    the reduced dim as its slow axis, so the common (trials first) case is never copied."""
    if isinstance(dims, int):
        dims = (dims,)
    nd = X.ndim
    dims = tuple(d if d >= 0 else d + nd for d in dims)
    if len(set(dims)) != len(dims):
        raise ValueError('Specified dimensions in `dims` must be unique for `reduce_` in `metrics.score()`.')
    if any(d < 0 or d >= nd for d in dims):
        raise ValueError('Specified dimensions in `dims` are out of range for `reduce_` in `metrics.score()`.')
    red = tuple(sorted(dims))
    keep = tuple(i for i in range(nd) if i not in red)
    Xm = X.permute(*keep, *red)
    return Xm.reshape(*Xm.shape[:len(keep)], prod(Xm.shape[len(keep):]))


def score(model, metric: Union[Metric, Tuple[Metric, ...]], X: torch.Tensor, y: Optional[torch.Tensor] = None
          ) -> Union[torch.Tensor, Dict[str, torch.Tensor]]:
    """Names in ``Metric.request`` resolve, in order, from the cache (``X``, ``y``), the metric's
    metadata (``y_b``), then the model (``Pipeline``: also its last step); callables are invoked on
    ``X`` once and shared by all metrics.  One metric -> tensor, several -> dict by name."""
    if isinstance(metric, Metric):
        metric = (metric,)
    cache = {'X': X, 'y': y}
    reduced: Dict[tuple, Dict[str, Optional[torch.Tensor]]] = {}
    out = {}
    for m in metric:
        red_key = m.reduce if isinstance(m.reduce, tuple) else (m.reduce,)
        args = []
        for r in m.request:
            if r not in cache:
                meta = m.metadata or {}          # reference defect 2: metadata may be None
                if r in meta:
                    cache[r] = meta[r]
                elif hasattr(model, r):
                    cache[r] = getattr(model, r, None)
                elif isinstance(model, Pipeline):
                    cache[r] = getattr(model[-1], r, None)
                else:
                    cache[r] = None
                if callable(cache[r]):
                    cache[r] = cache[r](X)
            slot = reduced.setdefault(red_key, {})
            if r not in slot:
                v = cache[r]
                slot[r] = reduce_(v, red_key) if v is not None else None
            args.append(slot[r])
        out[m.name] = m(*args)
    if len(metric) == 1:
        return out[metric[0].name]
    return out
