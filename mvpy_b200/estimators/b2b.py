"""Back-to-back regression -- drop-in for ``mvpy.estimators.B2B`` (mvpy/estimators/b2b.py:180-515).

A ridge decoder maps data to targets on the first half of the samples; on the second half its (optionally
standardised) predictions are regressed back onto the targets by a second ridge model, whose diagonal is the
causal contribution of each target (King et al. 2020).  Pure composition of hot-path pieces: both fits are the
LOO ridge kernels behind ``RidgeDecoder``, the standardisation is the ``Scaler`` kernels.

At the reference commit the constructor raises (it forwards ``normalise`` to ``RidgeCV``; SURVEY.md section 8c
defect 1, the same defect as ``RidgeDecoder``).  This class does what b2b.py:267-306 spells out, with
``normalise`` accepted and ignored."""
from __future__ import annotations

from typing import Union

import numpy as np
import sklearn.base
import torch

from ..preprocessing.scaler import _Scaler_torch
from .ridgedecoder import _RidgeDecoder_torch


class _B2B_torch(sklearn.base.BaseEstimator):
    """Attributes: ``alphas, fit_intercept, normalise, alpha_per_target, normalise_decoder``; after ``fit``:
    ``decoder_`` (X -> y on the first half), ``scaler_``, ``encoder_`` (decoded y -> y on the second half),
    ``causal_`` (f,) = diag(encoder_.coef_), ``pattern_`` (p, f) = decoder_.pattern_."""

    def __init__(self, alphas: torch.Tensor = torch.tensor([1]), fit_intercept: bool = True, normalise: bool = True,
                 alpha_per_target: bool = False, normalise_decoder: bool = True):
        self.alphas = alphas
        self.fit_intercept = fit_intercept
        self.normalise = normalise
        self.alpha_per_target = alpha_per_target
        self.normalise_decoder = normalise_decoder
        opts = dict(fit_intercept=fit_intercept, normalise=normalise, alpha_per_target=alpha_per_target)
        self.decoder_ = _RidgeDecoder_torch(alpha=alphas, **opts)
        self.encoder_ = _RidgeDecoder_torch(alpha=alphas, **opts)
        self.scaler_ = _Scaler_torch(with_mean=normalise_decoder, with_std=normalise_decoder)
        self.causal_ = None
        self.pattern_ = None

    def fit(self, X: torch.Tensor, y: torch.Tensor):
        if y.ndim == 1:
            y = y[:, None]
        half = X.shape[0] // 2
        self.decoder_.fit(X[:half], y[:half])
        decoded = self.decoder_.predict(X[half:])
        self.scaler_.fit(decoded)
        self.encoder_.fit(self.scaler_.transform(decoded), y[half:])
        self.causal_ = self.encoder_.coef_.diagonal()
        self.pattern_ = self.decoder_.pattern_.clone()
        return self

    def predict(self, X: torch.Tensor) -> torch.Tensor:
        if self.causal_ is None or self.pattern_ is None:
            raise ValueError('Estimator has not been fit yet.')
        return self.encoder_.predict(self.scaler_.transform(self.decoder_.predict(X)))

    def clone(self):
        return _B2B_torch(alphas=self.alphas, fit_intercept=self.fit_intercept, normalise=self.normalise,
                          alpha_per_target=self.alpha_per_target, normalise_decoder=self.normalise_decoder)


class B2B(sklearn.base.BaseEstimator):
    def __new__(cls, alphas: Union[torch.Tensor, np.ndarray, float, int] = 1, **kwargs):
        if isinstance(alphas, (float, int)):
            alphas = torch.tensor([alphas])
        if isinstance(alphas, list):
            alphas = torch.tensor(alphas)
        if isinstance(alphas, torch.Tensor):
            return _B2B_torch(alphas=alphas, **kwargs)
        raise ValueError(f'Alphas should be of type torch.Tensor (the B200 build has no numpy backend), but got {type(alphas)}.')
