"""``shapley_score`` -- same signature as mvpy/model_selection/shapley_score.py:12-152."""
from typing import Any, List, Optional, Tuple, Union

import torch

from .. import metrics
from ..crossvalidation import Validator
from .shapley import Shapley


def shapley_score(model, X: torch.Tensor, y: torch.Tensor, groups: Optional[Union[List, torch.Tensor]] = None,
                  dim: Optional[int] = None, cv: Union[int, Any] = 5,
                  metric: Optional[Union[metrics.Metric, Tuple[metrics.Metric]]] = None, return_shapley: bool = True,
                  n_permutations: int = 10, n_jobs: Optional[int] = None, n_jobs_validator: Optional[int] = None,
                  verbose: Union[int, bool] = False, verbose_validator: Union[int, bool] = False):
    validator = Validator(model, cv=cv, metric=metric, n_jobs=n_jobs_validator, verbose=verbose_validator)
    shapley = Shapley(validator, n_permutations=n_permutations, n_jobs=n_jobs, verbose=verbose).fit(X, y, groups=groups, dim=dim)
    if return_shapley:
        return shapley, shapley.score_
    return shapley.score_
