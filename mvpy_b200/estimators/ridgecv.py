"""Leave-one-out cross-validated ridge -- drop-in for ``mvpy.estimators.RidgeCV``
(mvpy/estimators/ridgecv.py:15-552).

The reference centres a copy of X, takes a thin SVD of [X, 1] (or an eigendecomposition of the dual
Gram when n <= p) and sweeps the alpha grid with the closed-form LOO error.  Here the same quantities
come from the centred Gram: ``mvb_trf_stats`` (tcgen05 split-TF32 contraction) -> ``mvb_ridge_solve``
(Jacobi eigensolve, per-row leverages, LOO losses, first-minimum alpha selection, coefficients), all
on the GPU; see SURVEY.md section 8 a-note for the equivalence."""
from __future__ import annotations

from typing import Dict, Optional, Tuple, Union

import numpy as np
import sklearn.base
import torch

from .. import _lib, metrics


def _as_alphas(alphas) -> torch.Tensor:
    if isinstance(alphas, (int, float)):
        return torch.tensor([alphas])
    if isinstance(alphas, list):
        return torch.tensor(alphas)
    return alphas


def solve_design(design: _lib.Design, y: torch.Tensor, alphas: torch.Tensor, fit_intercept: bool, alpha_per_target: bool,
                 stats=None, want_loss: bool = False, keep_gram: bool = True):
    """Shared fit core: y is (N, F, T) on the device of ``design``.  Returns dict with flat coef (F, p),
    intercept (F,), alpha (0-d or (F,)), best, loss, and the centred Gram (for patterns)."""
    if (stats is None and design.xp.dtype == torch.float32 and design.n_rows <= design.n_cols + int(fit_intercept)
            and design.n_cols <= 2048):
        # No more rows than unknowns: every training row is (nearly) interpolated, 1 - leverage is of order
        # alpha / eigenvalue and fp32 cannot form it (the reference's dual route for n <= p, ridgecv.py:135-190, never
        # does).  This regime is small by construction (the Gram is p x p with p >= n), so it runs on the fp64 kernels.
        wide = _lib.Design(design.xp.double(), design.n_time, design.n_lag, design.trial_idx, design.feat_idx)
        out = solve_design(wide, y.double(), alphas.double(), fit_intercept, alpha_per_target, None, want_loss, keep_gram)
        return {k: (v.float() if isinstance(v, torch.Tensor) and v.dtype == torch.float64 else v) for k, v in out.items()}
    if stats is None:
        stats = _lib.trf_stats(design, y, fit_intercept)
    G, XtY, mu, ybar = stats
    Gkeep = G.clone() if keep_gram else None          # the solve may overwrite G; patterns need cov(X) = G / (n - 1)
    coef, icpt, alpha, best, loss = _lib.ridge_solve(design, y, G, XtY, mu, ybar, alphas, fit_intercept, alpha_per_target,
                                                     want_loss=want_loss)
    return dict(coef=coef, intercept=icpt, alpha=alpha if alpha_per_target else alpha.reshape(()), best=best, loss=loss,
                gram=Gkeep, mu=mu, ybar=ybar)


class _RidgeCV_torch(sklearn.base.BaseEstimator):
    """Attributes after ``fit``: ``alpha_`` (0-d, or (n_targets,) with ``alpha_per_target``), ``coef_``
    (n_targets, n_features), ``intercept_`` ((1, n_targets) tensor, or the float 0.0 without intercept),
    ``metric_`` (default r2).  Inputs are never modified; outputs live on the input's device."""

    def __init__(self, alphas: Union[torch.Tensor, list, float, int] = 1, fit_intercept: bool = True,
                 alpha_per_target: bool = False):
        self.alphas = _as_alphas(alphas)
        self.fit_intercept = fit_intercept
        self.alpha_per_target = alpha_per_target
        self.alpha_ = None
        self.coef_ = None
        self.intercept_ = None
        self.loss_ = None
        self.metric_ = metrics.r2

    def fit(self, X: torch.Tensor, y: torch.Tensor):
        if X.shape[0] != y.shape[0]:
            raise ValueError(f'`X` and `y` must have the same number of samples, but got {X.shape[0]} and {y.shape[0]}')
        if y.ndim == 1:
            y = y[:, None]
        out_dev = X.device
        Xd = _lib.to_device(X)
        yd = _lib.to_device(y).to(Xd.dtype)
        self.alphas = self.alphas.to(X.dtype).to(X.device)            # ridgecv.py:125
        n, p = Xd.shape
        # dense design in "series" layout: one trial, p feature rows of n samples, a single lag
        xs = _lib.transpose(Xd.contiguous()).reshape(1, p, n)
        ys = _lib.transpose(yd.contiguous()).reshape(1, yd.shape[1], n)
        design = _lib.Design(xs, n_time=n, n_lag=1)
        fit = solve_design(design, ys, _lib.to_device(self.alphas).contiguous(), self.fit_intercept, self.alpha_per_target,
                           want_loss=True)
        self.coef_ = fit['coef'].to(out_dev)
        self.alpha_ = fit['alpha'].to(out_dev)
        self.loss_ = fit['loss'].to(out_dev)
        self.intercept_ = fit['intercept'][None, :].to(out_dev) if self.fit_intercept else 0.0
        self._gram = fit['gram']
        return self

    def predict(self, X: torch.Tensor) -> torch.Tensor:
        if (self.coef_ is None) or (self.intercept_ is None):
            raise ValueError('Model has not been fit yet.')
        out_dev = X.device
        Xd = _lib.to_device(X)
        n, p = Xd.shape
        xs = _lib.transpose(Xd.contiguous()).reshape(1, p, n)
        coef = _lib.to_device(self.coef_).to(Xd.dtype).contiguous()
        icpt = _lib.to_device(self.intercept_).to(Xd.dtype).reshape(-1).contiguous() if isinstance(self.intercept_, torch.Tensor) else None
        yh = _lib.trf_predict(_lib.Design(xs, n_time=n, n_lag=1), coef, icpt)      # (1, F, n)
        return _lib.transpose(yh[0]).to(out_dev)

    def score(self, X: torch.Tensor, y: torch.Tensor, metric: Optional[Union[metrics.Metric, Tuple[metrics.Metric]]] = None
              ) -> Union[torch.Tensor, Dict[str, torch.Tensor]]:
        if metric is None:
            metric = self.metric_
        return metrics.score(self, metric, X, y)

    def clone(self):
        return _RidgeCV_torch(alphas=self.alphas, fit_intercept=self.fit_intercept, alpha_per_target=self.alpha_per_target)


class RidgeCV(sklearn.base.BaseEstimator):
    """Dispatcher (ridgecv.py:370-552): tensor / scalar / list alphas give the CUDA estimator.  numpy
    alphas select sklearn's RidgeCV in the reference; this build serves the torch side only."""

    def __new__(cls, alphas: Union[torch.Tensor, np.ndarray, list, float, int] = 1, fit_intercept: bool = True,
                alpha_per_target: bool = False):
        alphas = _as_alphas(alphas)
        if isinstance(alphas, torch.Tensor):
            return _RidgeCV_torch(alphas=alphas, fit_intercept=fit_intercept, alpha_per_target=alpha_per_target)
        raise ValueError(f'Alphas should be of type torch.Tensor (the B200 build has no numpy backend), but got {type(alphas)}.')
