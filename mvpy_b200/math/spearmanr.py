"""Spearman correlation over the last dimension -- drop-in for ``mvpy.math.spearmanr`` / ``spearmanr_d``
(mvpy/math/spearmanr.py:13-90): Pearson correlation of the average ranks."""
from __future__ import annotations

from typing import Any

import torch

from .pearsonr import pearsonr
from .rank import rank


def spearmanr(x: torch.Tensor, y: torch.Tensor, *args: Any) -> torch.Tensor:
    if not (isinstance(x, torch.Tensor) and isinstance(y, torch.Tensor)):
        raise ValueError(f'`x` and `y` must be of the same type, but got `{type(x)}` and `{type(y)}` instead.')
    if x.shape != y.shape:
        raise ValueError('`x` and `y` must have the same shape.')
    return pearsonr(rank(x), rank(y))


def spearmanr_d(x: torch.Tensor, y: torch.Tensor, *args: Any) -> torch.Tensor:
    return 1 - spearmanr(x, y)
