"""Pearson metric instance (mvpy/metrics/pearsonr.py:11-50): requests ('y', 'predict'), reduces dim 0."""
from dataclasses import dataclass
from typing import Callable, Tuple, Union

from .metric import Metric
from ..math import pearsonr as _pearsonr


@dataclass
class Pearsonr(Metric):
    name: str = 'pearsonr'
    request: Tuple[str, ...] = ('y', 'predict')
    reduce: Union[int, Tuple[int, ...]] = (0,)
    f: Callable = _pearsonr


pearsonr = Pearsonr()
