"""Pearson correlation over the last dimension -- drop-in for ``mvpy.math.pearsonr``
(mvpy/math/pearsonr.py:35-128)."""
from __future__ import annotations

from typing import Any

import torch

from ._reduce import run_score


def pearsonr(x: torch.Tensor, y: torch.Tensor, *args: Any) -> torch.Tensor:
    """Centred dot product over sqrt of the centred sums of squares, last dim; NaN on zero variance
    like the reference (mvpy/math/pearsonr.py:55-56).  One fused CUDA kernel (``mvb_score``)."""
    if not (isinstance(x, torch.Tensor) and isinstance(y, torch.Tensor)):
        raise ValueError(f'`x` and `y` must be of the same type, but got `{type(x)}` and `{type(y)}` instead.')
    if x.shape != y.shape:
        raise ValueError('`x` and `y` must have the same shape.')
    return run_score(x, y, None, False, True)[1]
