"""Accuracy metric instance (mvpy/metrics/accuracy.py:11-52): requests ('y', 'predict'), reduces dim 0."""
from dataclasses import dataclass
from typing import Callable, Tuple, Union

from .metric import Metric
from ..math import accuracy as _accuracy


@dataclass
class Accuracy(Metric):
    name: str = 'accuracy'
    request: Tuple[str, ...] = ('y', 'predict')
    reduce: Union[int, Tuple[int, ...]] = (0,)
    f: Callable = _accuracy


accuracy = Accuracy()
