"""``cross_val_score`` -- same signature as mvpy/crossvalidation/cross_val_score.py:16-99."""
from __future__ import annotations

from typing import Any, Optional, Tuple, Union

import torch

from ..metrics import Metric
from .validator import Validator


def cross_val_score(model, X: torch.Tensor, y: Optional[torch.Tensor] = None, groups: Optional[torch.Tensor] = None,
                    cv: Optional[Union[int, Any]] = 5, metric: Optional[Union[Metric, Tuple[Metric]]] = None,
                    return_validator: bool = True, n_jobs: Optional[int] = None, verbose: Union[int, bool] = False):
    validator = Validator(model=model, cv=cv, metric=metric, n_jobs=n_jobs, verbose=verbose).fit(X, y, groups=groups)
    if return_validator:
        return validator, validator.score_
    return validator.score_
