"""R^2 over the last dimension -- drop-in for ``mvpy.math.r2`` (mvpy/math/r2.py:57-147)."""
from __future__ import annotations

from typing import Optional

import torch

from ._reduce import run_score


def r2(y: torch.Tensor, y_h: torch.Tensor, y_b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``1 - sum((y - y_h)^2) / sum((y - y_b)^2)`` over the last dim; ``y_b`` defaults to the mean of
    ``y``; cells whose denominator is not positive score 0 (mvpy/math/r2.py:80-100).  Runs as one fused
    CUDA kernel (``mvb_score``); the result lives on the device of ``y``."""
    if not (isinstance(y, torch.Tensor) and isinstance(y_h, torch.Tensor) and (y_b is None or isinstance(y_b, torch.Tensor))):
        raise ValueError(f'`y`, `y_h` and, if supplied, `y_b` must be of type torch.Tensor, but received '
                         f'{type(y)}, {type(y_h)} and {type(y_b)} instead.')
    if y.shape != y_h.shape:
        raise ValueError(f'`y` and `y_h` must have the same shape, but got {y.shape} and {y_h.shape}.')
    if y_b is not None and y_b.ndim != y.ndim:
        raise ValueError(f'When supplying `y_b`, it must have the same number of dimensions as `y`, '
                         f'but got {y.ndim} and {y_b.ndim}.')
    return run_score(y, y_h, y_b, True, False)[0]
