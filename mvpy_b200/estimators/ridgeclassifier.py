"""Ridge classifier -- drop-in for ``mvpy.estimators.RidgeClassifier`` (mvpy/estimators/ridgeclassifier.py:266-786).

Classification as LOO ridge regression onto -1 / +1 one-hot targets (one column per class of every label feature):
the fit is the ``RidgeDecoder`` CUDA path on the binarised labels, the decision values are its predictions, the
predicted label of a feature is its class with the largest decision value.

At the reference commit the class cannot be constructed (it forwards ``normalise`` to ``RidgeCV`` through
``RidgeDecoder``; SURVEY.md section 8c defect 1) and its ``clone`` reads a misspelt attribute; this class does what
ridgeclassifier.py:356-515 spells out, with ``normalise`` accepted and ignored."""
from __future__ import annotations

from typing import Dict, Optional, Tuple, Union

import numpy as np
import sklearn.base
import torch

from .. import metrics
from ..preprocessing.labelbinariser import _LabelBinariser_torch
from .ridgedecoder import _RidgeDecoder_torch


class _RidgeClassifier_torch(sklearn.base.BaseEstimator):
    """Attributes: ``alpha, fit_intercept, normalise, alpha_per_target, estimator`` (the ridge decoder),
    ``binariser_``; after ``fit``: ``coef_`` (n_classes, n_channels), ``intercept_``, ``pattern_``; ``metric_`` = accuracy."""

    def __init__(self, alpha: torch.Tensor, fit_intercept: bool = True, normalise: bool = True, alpha_per_target: bool = False):
        self.alpha = alpha
        self.fit_intercept = fit_intercept
        self.normalise = normalise
        self.alpha_per_target = alpha_per_target
        self.estimator = _RidgeDecoder_torch(self.alpha, fit_intercept=fit_intercept, normalise=normalise, alpha_per_target=alpha_per_target)
        self.binariser_ = _LabelBinariser_torch(neg_label=-1.0, pos_label=1.0)
        self.intercept_ = None
        self.coef_ = None
        self.pattern_ = None
        self.metric_ = metrics.accuracy

    def _blocks(self):
        start = 0
        for labels in self.binariser_.labels_:
            yield start, start + int(labels.shape[0])
            start += int(labels.shape[0])

    def fit(self, X: torch.Tensor, y: torch.Tensor):
        if len(X.shape) != 2:
            raise ValueError(f'X must be of shape (n_samples, n_features), but got {X.shape}.')
        if len(y.shape) == 1:
            y = y[:, None]
        L = self.binariser_.fit_transform(y)
        self.estimator.fit(X, L.to(X.dtype))
        self.coef_ = self.estimator.coef_
        self.intercept_ = self.estimator.intercept_
        self.pattern_ = self.estimator.pattern_
        return self

    def decision_function(self, X: torch.Tensor) -> torch.Tensor:
        if self.coef_ is None or self.intercept_ is None or self.pattern_ is None:
            raise ValueError('The RidgeClassifier has not been fitted yet.')
        return self.estimator.predict(X)

    def predict(self, X: torch.Tensor) -> torch.Tensor:
        df = self.decision_function(X)
        onehot = torch.zeros_like(df)
        rows = torch.arange(df.shape[0], device=df.device)
        for s, e in self._blocks():
            onehot[rows, s + torch.argmax(df[:, s:e], dim=1)] = 1.0
        return self.binariser_.inverse_transform(onehot)

    def predict_proba(self, X: torch.Tensor) -> torch.Tensor:
        """Logistic squashing of the decision values, normalised within each label feature (uncalibrated)."""
        p = 1 / (1 + torch.exp(-self.decision_function(X)))
        for s, e in self._blocks():
            p[..., s:e] = p[..., s:e] / p[..., s:e].sum(-1, keepdim=True)
        return p

    def score(self, X: torch.Tensor, y: torch.Tensor, metric: Optional[Union[metrics.Metric, Tuple[metrics.Metric]]] = None
              ) -> Union[torch.Tensor, Dict[str, torch.Tensor]]:
        if metric is None:
            metric = self.metric_
        return metrics.score(self, metric, X, y)

    def clone(self):
        return _RidgeClassifier_torch(self.alpha, fit_intercept=self.fit_intercept, normalise=self.normalise,
                                      alpha_per_target=self.alpha_per_target)


class RidgeClassifier(sklearn.base.BaseEstimator):
    """Dispatcher (ridgeclassifier.py:586-651): the ridge classifier inside the OvR / OvO ``Classifier`` wrapper."""

    def __new__(cls, alphas: Union[torch.Tensor, np.ndarray, float, int] = 1, method: str = 'OvR', fit_intercept: bool = True,
                normalise: bool = True, alpha_per_target: bool = False):
        from .classifier import _Classifier_torch
        if isinstance(alphas, (float, int)):
            alphas = torch.tensor([alphas])
        if isinstance(alphas, list):
            alphas = torch.tensor(alphas)
        if isinstance(alphas, torch.Tensor):
            return _Classifier_torch(_RidgeClassifier_torch, method=method, arguments=[alphas],
                                     kwarguments=dict(fit_intercept=fit_intercept, normalise=normalise, alpha_per_target=alpha_per_target))
        raise TypeError('`alphas` must be of type torch.Tensor, float or int (the B200 build has no numpy backend), '
                        f'but got {type(alphas)}.')
