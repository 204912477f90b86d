"""Average ranks along the last dimension -- drop-in for ``mvpy.math.rank`` (mvpy/math/rank.py:62-177).
Ties receive the mean of the positions they occupy (1-based), e.g. ``rank([2, .5, 1, 1]) = [4, 1, 2.5, 2.5]``.
Runs as one CUDA kernel (``mvb_rank``) on the (R, K) view the metrics use; integer input is ranked in fp64
like the reference."""
from __future__ import annotations

import torch

from .. import _lib
from ._reduce import as_rows


def rank(x: torch.Tensor) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        raise ValueError(f'`x` must be of type torch.Tensor, but got {type(x)}.')
    out_dev = x.device
    if not torch.is_floating_point(x):
        x = x.to(torch.float64)
    xd = _lib.to_device(x)
    if xd.dtype not in (torch.float32, torch.float64):
        xd = xd.to(torch.float32)
    if xd.numel() == 0:
        return xd.clone().to(out_dev)
    rows = as_rows(xd)                                   # (R, K), R = the ranked (last) dim
    r = _lib.rank_rows(rows)
    return r.reshape((xd.shape[-1],) + tuple(xd.shape[:-1])).movedim(0, -1).to(out_dev)
