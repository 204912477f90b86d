"""``metrics.score`` -- resolve what each metric requests, share one prediction, reduce, evaluate.
Behaviour follows mvpy/metrics/score.py:17-122 (lookup order, caching per reduce tuple, one metric -> tensor)."""
from __future__ import annotations

from math import prod
from typing import Dict, Optional, Tuple, Union

import torch
from sklearn.pipeline import Pipeline

from .metric import Metric


def _normalise_dims(dims: Union[int, Tuple[int, ...]], nd: int) -> Tuple[int, ...]:
    dims = (dims,) if isinstance(dims, int) else tuple(dims)
    pos = tuple(d + nd if d < 0 else d for d in dims)
    if len(set(pos)) != len(pos):
        raise ValueError('Specified dimensions in `dims` must be unique for `reduce_` in `metrics.score()`.')
    if not all(0 <= d < nd for d in pos):
        raise ValueError('Specified dimensions in `dims` are out of range for `reduce_` in `metrics.score()`.')
    return tuple(sorted(pos))


def reduce_(X: torch.Tensor, dims: Union[int, Tuple[int, ...]]) -> torch.Tensor:
    """Put ``dims`` (ascending) behind the kept dims and flatten them into one trailing dim
    (mvpy/metrics/score.py:17-46).  Stays a view whenever torch allows it; the scoring kernel reads
    the reduced dim as its slow axis, so the common (trials first) case is never copied."""
    red = _normalise_dims(dims, X.ndim)
    keep = [i for i in range(X.ndim) if i not in red]
    moved = X.permute(*keep, *red)
    return moved.reshape(*moved.shape[:len(keep)], prod(moved.shape[len(keep):]))


class _Requests:
    """Values a metric may ask for, by name.  Lookup order: the call's ``X`` / ``y``, the metric's metadata
    (``y_b``), the model, and for a ``Pipeline`` its last step; callables are invoked on ``X`` exactly once
    (so all metrics share one prediction).  Reduced views are memoised per reduce tuple."""

    def __init__(self, model, X, y):
        self.model, self.X = model, X
        self.raw = {'X': X, 'y': y}
        self.views: Dict[tuple, Dict[str, Optional[torch.Tensor]]] = {}

    def _lookup(self, name: str, metric: Metric):
        meta = metric.metadata or {}                 # metadata may be None on an ungranted metric (defect 2)
        if name in meta:
            return meta[name]
        if hasattr(self.model, name):
            return getattr(self.model, name, None)
        if isinstance(self.model, Pipeline):
            return getattr(self.model[-1], name, None)
        return None

    def value(self, name: str, metric: Metric):
        if name not in self.raw:
            found = self._lookup(name, metric)
            self.raw[name] = found(self.X) if callable(found) else found
        return self.raw[name]

    def reduced(self, name: str, metric: Metric):
        key = metric.reduce if isinstance(metric.reduce, tuple) else (metric.reduce,)
        per_key = self.views.setdefault(key, {})
        if name not in per_key:
            v = self.value(name, metric)
            per_key[name] = None if v is None else reduce_(v, key)
        return per_key[name]


def score(model, metric: Union[Metric, Tuple[Metric, ...]], X: torch.Tensor, y: Optional[torch.Tensor] = None
          ) -> Union[torch.Tensor, Dict[str, torch.Tensor]]:
    """One metric -> tensor, several -> dict keyed by ``Metric.name``."""
    metrics = (metric,) if isinstance(metric, Metric) else tuple(metric)
    req = _Requests(model, X, y)
    results = {m.name: m(*[req.reduced(r, m) for r in m.request]) for m in metrics}
    return results[metrics[0].name] if len(metrics) == 1 else results
