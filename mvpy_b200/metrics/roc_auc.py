"""ROC-AUC metric instance (mvpy/metrics/roc_auc.py:11-112): requests ('y', 'decision_function'), reduces dim 0.

The decision values of a classifier hold one column per class of every label feature, features side by side; the call
splits them per feature (two classes: the last column is the score; more: one-versus-rest over all of them) and returns
(n_features[, n_timepoints])."""
from dataclasses import dataclass
from typing import Callable, Tuple, Union

import torch

from .metric import Metric
from ..math import roc_auc as _roc_auc


@dataclass
class Roc_auc(Metric):
    name: str = 'roc_auc'
    request: Tuple[str, ...] = ('y', 'decision_function')
    reduce: Union[int, Tuple[int, ...]] = (0,)
    f: Callable = _roc_auc

    def __call__(self, y: torch.Tensor, df: torch.Tensor) -> torch.Tensor:
        if y.ndim > 2:                                   # (..., features, time, samples) -> (..., time, features, samples)
            y, df = y.swapaxes(-2, -3), df.swapaxes(-2, -3)
        n_classes = [int(torch.unique(y[..., i, :]).shape[0]) for i in range(y.shape[-2])]
        offsets = [0] + n_classes[:-1]                   # as the reference has it (roc_auc.py:91): the previous feature's
        scores = []                                      # class count, not a running sum
        for i, k in enumerate(n_classes):
            block = df[..., offsets[i]:offsets[i] + k, :]
            scores.append(self.f(y[..., i, :], block if block.shape[-2] > 2 else block[..., -1, :]))
        return torch.stack(scores, dim=0)


roc_auc = Roc_auc()
