"""``Metric`` request/reduce protocol -- same contract as mvpy/metrics/metric.py:12-183."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, Optional, Tuple, Union


@dataclass
class Metric:
    """name: key in multi-metric outputs; request: names resolved by ``metrics.score`` (``'y'``,
    ``'X'``, metadata keys such as ``'y_b'``, or model attributes/methods such as ``'predict'``);
    reduce: dims scored over (moved last and flattened); f: callable on the resolved tensors;
    metadata: values granted by the validator (the training-set baseline ``y_b``)."""
    name: str = 'metric'
    request: Tuple[str, ...] = ('y', 'predict')
    reduce: Union[int, Tuple[int, ...]] = (0,)
    f: Callable = lambda x: x
    metadata: Optional[Dict[str, Any]] = None

    def __call__(self, *args: Any, **kwargs: Any):
        return self.f(*args, **kwargs)

    def grant(self, metadata: Dict[str, Any]) -> None:
        if not isinstance(metadata, dict):
            raise ValueError(f'When granting metadata, metadata must be a dictionary of `key` => `value` pairs, '
                             f'but got {type(metadata)}.')
        if self.metadata is None:
            self.metadata = dict(metadata)
        else:
            self.metadata.update(metadata)

    def mutate(self, name: Optional[str] = None, request=None, reduce=None, f: Optional[Callable] = None,
               metadata: Optional[Dict[str, Any]] = None) -> 'Metric':
        return Metric(name=name or self.name, request=request or self.request, reduce=reduce or self.reduce,
                      f=f or self.f, metadata=metadata or self.metadata)
