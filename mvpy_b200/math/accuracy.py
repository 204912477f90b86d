"""Accuracy over the last dimension -- drop-in for ``mvpy.math.accuracy`` (mvpy/math/accuracy.py:13-91):
the fraction of positions where ``x`` and ``y`` are equal.  One CUDA kernel (``mvb_accuracy``) on the (R, K) view the
metrics use; the result has ``x``'s dtype like the reference (floating inputs; labels are floats on this path)."""
from __future__ import annotations

import torch

from .. import _lib
from ._reduce import as_rows


def accuracy(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    if not (isinstance(x, torch.Tensor) and isinstance(y, torch.Tensor)):
        raise ValueError(f'`x` and `y` must be of the same type, but got `{type(x)}` and `{type(y)}` instead.')
    if x.shape != y.shape:
        raise ValueError('`x` and `y` must have the same shape.')
    out_dev = x.device
    out_dtype = x.dtype if torch.is_floating_point(x) else torch.float32      # (the reference's mean() rejects integer input)
    work = x.dtype if x.dtype in (torch.float32, torch.float64) else torch.float64        # integer labels compare exactly in fp64
    xd, yd = _lib.to_device(x).to(work), _lib.to_device(y).to(work)
    acc = _lib.accuracy_rows(as_rows(xd), as_rows(yd)).reshape(x.shape[:-1])
    return acc.to(out_dtype).to(out_dev)
