"""Spearman metric instance (mvpy/metrics/spearmanr.py:11-50): requests ('y', 'predict'), reduces dim 0."""
from dataclasses import dataclass
from typing import Callable, Tuple, Union

from .metric import Metric
from ..math import spearmanr as _spearmanr


@dataclass
class Spearmanr(Metric):
    name: str = 'spearmanr'
    request: Tuple[str, ...] = ('y', 'predict')
    reduce: Union[int, Tuple[int, ...]] = (0,)
    f: Callable = _spearmanr


spearmanr = Spearmanr()
