"""Ridge encoder -- drop-in for ``mvpy.estimators.RidgeEncoder`` (mvpy/estimators/ridgeencoder.py:246-613).

2-D input is plain RidgeCV.  3-D input (trials, features, time) fits ONE ridge with a shared alpha and
intercept on the block design X_t[i*T + j, j*F:(j+1)*F] = X[i, :, j] (ridgeencoder.py:292-304, built there
with a Python double loop); the design is scattered on the device in one indexing op and solved by
the same CUDA path."""
from __future__ import annotations

from typing import Dict, Optional, Tuple, Union

import numpy as np
import sklearn.base
import torch

from .. import _lib, metrics
from .ridgecv import _RidgeCV_torch


class _RidgeEncoder_torch(sklearn.base.BaseEstimator):
    """Attributes: ``alphas, kwargs, estimator, expanded_, intercept_, coef_`` ((d, f) or (d, f, t)),
    ``n_, f_, t_, d_, metric_``."""

    def __init__(self, alphas: torch.Tensor = torch.tensor([1.0]), **kwargs):
        self.alphas = alphas
        self.kwargs = kwargs
        self.estimator = _RidgeCV_torch(alphas=self.alphas, **self.kwargs)
        self.expanded_ = None
        self.intercept_ = None
        self.coef_ = None
        self.n_ = None
        self.f_ = None
        self.t_ = None
        self.d_ = None
        self.metric_ = metrics.r2

    @staticmethod
    def _expand(X: torch.Tensor) -> torch.Tensor:
        N, F, T = X.shape
        out = torch.zeros((N, T, T, F), dtype=X.dtype, device=X.device)
        idx = torch.arange(T, device=X.device)
        out[:, idx, idx, :] = X.permute(0, 2, 1)
        return out.reshape(N * T, T * F)

    def fit(self, X: torch.Tensor, y: torch.Tensor):
        if y.ndim == 1:
            y = y[:, None]
        if (X.ndim == 3) & (y.ndim == 3):
            self.expanded_ = True
            self.n_, self.f_, self.t_ = X.shape
            self.d_ = y.shape[1]
            Xd = _lib.to_device(X)
            X2 = self._expand(Xd)
            y2 = _lib.to_device(y).permute(0, 2, 1).reshape(-1, y.shape[1])
            out_dev = X.device
            self.estimator.fit(X2, y2)
            est = self.estimator
            est.coef_, est.alpha_ = est.coef_.to(out_dev), est.alpha_.to(out_dev)
            if isinstance(est.intercept_, torch.Tensor):
                est.intercept_ = est.intercept_.to(out_dev)
            est.alphas = est.alphas.to(out_dev)
        else:
            self.estimator.fit(X, y)
        self.intercept_ = self.estimator.intercept_.clone() if isinstance(self.estimator.intercept_, torch.Tensor) else self.estimator.intercept_
        self.coef_ = self.estimator.coef_.clone()
        if self.expanded_:
            self.coef_ = self.coef_.reshape((self.d_, self.t_, self.f_)).swapaxes(1, 2)
        return self

    def predict(self, X: torch.Tensor) -> torch.Tensor:
        if (self.intercept_ is None) | (self.coef_ is None):
            raise ValueError('Encoder has not been fitted yet.')
        dims = X.shape
        if self.expanded_:
            out_dev = X.device
            y = self.estimator.predict(self._expand(_lib.to_device(X))).to(out_dev)
            return y.reshape((dims[0], dims[-1], self.d_)).swapaxes(1, 2)
        return self.estimator.predict(X)

    def score(self, X: torch.Tensor, y: torch.Tensor, metric: Optional[Union[metrics.Metric, Tuple[metrics.Metric]]] = None
              ) -> Union[torch.Tensor, Dict[str, torch.Tensor]]:
        if metric is None:
            metric = self.metric_
        return metrics.score(self, metric, X, y)

    def clone(self):
        return _RidgeEncoder_torch(alphas=self.alphas, **self.kwargs)


class RidgeEncoder(sklearn.base.BaseEstimator):
    def __new__(cls, alphas: Union[torch.Tensor, np.ndarray, float, int] = 1, **kwargs):
        if isinstance(alphas, (float, int)):
            alphas = torch.tensor([alphas])
        if isinstance(alphas, list):
            alphas = torch.tensor(alphas)
        if isinstance(alphas, torch.Tensor):
            return _RidgeEncoder_torch(alphas=alphas, **kwargs)
        raise ValueError(f'Alphas should be of type torch.Tensor (the B200 build has no numpy backend), but got {type(alphas)}.')
