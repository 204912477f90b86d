"""Ridge decoder with Haufe patterns -- drop-in for ``mvpy.estimators.RidgeDecoder``
(mvpy/estimators/ridgedecoder.py:222-528).

At the reference commit the public constructor raises (it forwards a ``normalise`` keyword that
``RidgeCV`` does not accept; SURVEY.md section 8c defect 1).  This class does what that code evidently
intends: LOO ridge with per-target alphas by default, then ``pattern_ = cov(X) coef^T pinv(cov(y))``
(ridgedecoder.py:244-285); ``normalise`` is accepted and ignored."""
from __future__ import annotations

from typing import Dict, Optional, Tuple, Union

import numpy as np
import sklearn.base
import torch

from .. import _lib, metrics
from .ridgecv import _RidgeCV_torch


class _RidgeDecoder_torch(sklearn.base.BaseEstimator):
    """Attributes: ``alpha, fit_intercept, normalise, alpha_per_target, estimator, pattern_`` (p, f),
    ``coef_`` (f, p), ``intercept_``, ``alpha_``, ``metric_``."""

    def __init__(self, alpha: torch.Tensor, **kwargs):
        self.alpha = alpha
        self.fit_intercept = kwargs.get('fit_intercept', True)
        self.normalise = kwargs.get('normalise', True)
        self.alpha_per_target = kwargs.get('alpha_per_target', True)
        self.estimator = _RidgeCV_torch(alphas=alpha, fit_intercept=self.fit_intercept, alpha_per_target=self.alpha_per_target)
        self.pattern_ = None
        self.coef_ = None
        self.intercept_ = None
        self.alpha_ = None
        self.metric_ = metrics.r2

    def fit(self, X: torch.Tensor, y: torch.Tensor):
        if X.shape[0] != y.shape[0]:
            raise ValueError('X and y must have the same number of samples.')
        if y.ndim == 1:
            y = y[:, None]
        est = self.estimator.fit(X, y)
        self.coef_, self.intercept_, self.alpha_ = est.coef_, est.intercept_, est.alpha_
        # cov(X) is the centred Gram the fit already built (always demeaned for patterns)
        n = X.shape[0]
        out_dev = X.device
        if self.fit_intercept and est._gram is not None:
            Sx = est._gram / float(n - 1)
        else:
            Xd = _lib.to_device(X)
            xs = _lib.transpose(Xd.contiguous()).reshape(1, Xd.shape[1], n)
            yd = _lib.transpose(_lib.to_device(y).to(Xd.dtype).contiguous()).reshape(1, y.shape[1], n)
            Sx = _lib.trf_stats(_lib.shift_design(_lib.Design(xs, n_time=n, n_lag=1))[0], yd, True)[0] / float(n - 1)
        coef = _lib.to_device(self.coef_).to(Sx.dtype)
        if y.shape[1] > 1:
            yd = _lib.to_device(y).to(Sx.dtype)
            Py = _lib.pinv_symmetric(_lib.covariance(yd))      # pseudo-inverse of the tiny (f x f) target covariance
            self.pattern_ = _lib.matmul_nt(_lib.matmul_nt(Sx, coef), Py.T.contiguous()).to(out_dev)
        else:
            self.pattern_ = (_lib.matmul_nt(Sx, coef) * (1.0 / float(n - 1))).to(out_dev)
        return self

    def predict(self, X: torch.Tensor) -> torch.Tensor:
        if self.pattern_ is None:
            raise ValueError('The decoder has not been fitted yet.')
        if X.shape[1] != self.estimator.coef_.shape[1]:
            raise ValueError('X and ß must have the same number of features.')
        return self.estimator.predict(X)

    def score(self, X: torch.Tensor, y: torch.Tensor, metric: Optional[Union[metrics.Metric, Tuple[metrics.Metric]]] = None
              ) -> Union[torch.Tensor, Dict[str, torch.Tensor]]:
        if metric is None:
            metric = self.metric_
        return metrics.score(self, metric, X, y)

    def clone(self):
        return _RidgeDecoder_torch(alpha=self.alpha, fit_intercept=self.fit_intercept, normalise=self.normalise,
                                   alpha_per_target=self.alpha_per_target)


class RidgeDecoder(sklearn.base.BaseEstimator):
    def __new__(cls, alphas: Union[torch.Tensor, np.ndarray, float, int] = 1, **kwargs):
        if isinstance(alphas, (float, int)):
            alphas = torch.tensor([alphas])
        if isinstance(alphas, list):
            alphas = torch.tensor(alphas)
        if isinstance(alphas, torch.Tensor):
            return _RidgeDecoder_torch(alpha=alphas, **kwargs)
        raise ValueError(f'Alphas should be of type torch.Tensor (the B200 build has no numpy backend), but got {type(alphas)}.')
