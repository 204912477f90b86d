"""Area under the ROC curve over the last dimension -- drop-in for ``mvpy.math.roc_auc`` (mvpy/math/roc_auc.py:84-212).

Binary labels: the larger label is the positive class; a (…, 2, n) score keeps its second row.  More than two labels:
one-versus-rest per class against the matching score row, macro-averaged ignoring undefined classes.  The reference
ranks the scores and uses the rank-sum formula; ``mvb_roc_auc`` counts the same (positive, negative) pairs directly,
ties as one half, NaN where a cell has a single class."""
from __future__ import annotations

import torch

from .. import _lib
from ._reduce import as_rows


def roc_auc(y_true: torch.Tensor, y_score: torch.Tensor) -> torch.Tensor:
    if not (isinstance(y_true, torch.Tensor) and isinstance(y_score, torch.Tensor)):
        raise ValueError(f'`y_true` and `y_score` must be of the same type, but got `{type(y_true)}` and `{type(y_score)}` instead.')
    labels = torch.unique(y_true)
    if labels.shape[0] < 2:
        raise ValueError('`y_true` must have at least two unique values.')
    multiclass = labels.shape[0] > 2
    if multiclass:
        # (..., n) labels -> (..., n_classes, n) one-versus-rest indicators, class order = sorted labels
        positive = (y_true.unsqueeze(-2) == labels.reshape((-1, 1)).to(y_true.device))
        if y_score.ndim < 2 or y_score.shape[-2] != positive.shape[-2]:
            raise ValueError(f'For multiclass to work, `y_score` must have the same number of dimensions as there are labels in '
                             f'`y_true`, but got {y_score.shape[-2] if y_score.ndim > 1 else 1} and {positive.shape[-2]}.')
    else:
        if y_score.ndim > 2 and y_score.shape[-2] == 2:
            y_score = y_score[..., 1, :]
        positive = y_true == labels.max()
    if y_score.shape != positive.shape:
        raise ValueError(f'`y_true` and `y_score` must have the same shape, but got {tuple(positive.shape)} and {tuple(y_score.shape)}.')
    out_dev = y_score.device
    sd = _lib.to_device(y_score)
    if sd.dtype not in (torch.float32, torch.float64):
        sd = sd.to(torch.float32)
    pd = _lib.to_device(positive).to(sd.dtype)
    auc = _lib.roc_auc_rows(as_rows(pd), as_rows(sd)).reshape(sd.shape[:-1])
    if multiclass:
        auc = torch.nanmean(auc, -1)
    return auc.to(out_dev)
