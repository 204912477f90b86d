"""Shared plumbing for the last-dim metrics: present (..., R) tensors to the scoring kernel as [R][K]."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import _lib


def as_rows(x: torch.Tensor) -> torch.Tensor:
    """(..., R) -> contiguous (R, K).  ``metrics.score`` hands over views whose reduced dim was moved
    last; when that view is just a permutation of a contiguous (R, ...) tensor no copy is made."""
    xr = x.movedim(-1, 0)
    return xr.reshape(xr.shape[0], -1) if xr.is_contiguous() else xr.contiguous().reshape(xr.shape[0], -1)


def run_score(y: torch.Tensor, y_h: torch.Tensor, y_b: Optional[torch.Tensor], want_r2: bool, want_pearson: bool
              ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    out_dev = y.device
    shape = y.shape[:-1]
    yd, yhd = _lib.to_device(y), _lib.to_device(y_h)
    if yhd.dtype != yd.dtype:
        yhd = yhd.to(yd.dtype)
    y2, yh2 = as_rows(yd), as_rows(yhd)
    if y2.shape[0] == 0 or y2.shape[1] == 0:
        empty = torch.zeros(shape, dtype=y.dtype, device=out_dev)
        return (empty if want_r2 else None), (empty.clone() if want_pearson else None)
    yb1 = None
    if y_b is not None:
        yb1 = _lib.to_device(y_b).to(yd.dtype).expand(*shape, 1).reshape(-1).contiguous()
    r2, pr = _lib.score(y2, yh2, yb1, want_r2, want_pearson)
    r2 = None if r2 is None else r2.reshape(shape).to(out_dev)
    pr = None if pr is None else pr.reshape(shape).to(out_dev)
    return r2, pr
