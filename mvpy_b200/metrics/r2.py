"""R2 metric instance (mvpy/metrics/r2.py:11-53): requests ('y', 'predict', 'y_b'), reduces dim 0."""
from dataclasses import dataclass
from typing import Any, Callable, Dict, Optional, Tuple, Union

from .metric import Metric
from ..math import r2 as _r2


@dataclass
class R2(Metric):
    name: str = 'r2'
    request: Tuple[str, ...] = ('y', 'predict', 'y_b')
    reduce: Union[int, Tuple[int, ...]] = (0,)
    f: Callable = _r2
    metadata: Optional[Dict[str, Any]] = None


r2 = R2()
