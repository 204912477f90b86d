"""Time-delayed ridge (mTRF / stimulus reconstruction) -- drop-in for ``mvpy.estimators.TimeDelayed``
(mvpy/estimators/timedelayed.py:318-817).

The reference gathers every lag into a (trials*time, channels*lags) matrix and hands it to RidgeCV.
Here the stimulus is only edge-padded (``mvb_trf_pad``): X^T X is built from windows of the padded series (TMA +
overlapping UMMA descriptors, ``toeplitz_gram.cuh``), X^T y, X Q^T, X beta and the prediction read it through offset
tables, so the lag expansion never exists in HBM -- neither as a matrix nor as pre-split operand planes."""
from __future__ import annotations

import os
import threading
from typing import Dict, Optional, Tuple, Union

import numpy as np
import sklearn.base
import torch

from .. import _lib, metrics
from .ridgecv import _RidgeCV_torch, solve_design


SHARED_FOLD_ENTRIES = 8      # per-driver cache of per-fold full-predictor statistics (p x p Gram each): bounded
_SHARED_LOCK = threading.Lock()


def lag_window(t_min: float, t_max: float, fs: float) -> np.ndarray:
    """Integer lag window; timedelayed.py:376-380 passes both bounds through abs(ceil(.)), so the window
    always runs from -|ceil(t_min fs)| to +|ceil(t_max fs)| and contains 0."""
    t_neg = np.arange(np.abs(np.ceil(t_min * fs)) + 1)
    t_pos = np.arange(np.abs(np.ceil(t_max * fs)) + 1)
    return np.unique(np.concatenate((-t_neg[::-1], t_pos)))


class _TimeDelayed_torch(sklearn.base.BaseEstimator):
    """Attributes: ``alphas, kwargs, patterns, t_min, t_max, fs, window`` (int32 lags), ``estimator`` (the
    inner RidgeCV holding ``alpha_``, flat ``coef_`` (f, c*w), ``intercept_``), ``f_, c_, w_``,
    ``intercept_`` (1, f), ``coef_`` (f, c, w) stored lag-reversed, ``pattern_``, ``metric_``."""

    _mvb_fold_lanes = True       # a Validator may enqueue this estimator's folds on concurrent stream lanes (_streams.py)

    def __init__(self, t_min: float, t_max: float, fs: int, alphas: torch.Tensor = torch.tensor([1]), patterns: bool = False,
                 **kwargs):
        self.alphas = alphas
        self.kwargs = kwargs
        self.patterns = patterns
        self.t_min, self.t_max = t_min, t_max
        self.fs = fs
        win = lag_window(t_min, t_max, fs)
        self.window = torch.from_numpy(win).to(torch.int32).to(self.alphas.device)
        self._lag_min, self._lag_max = int(win.min()), int(win.max())
        self.estimator = _RidgeCV_torch(alphas=self.alphas, **self.kwargs)
        self.f_ = None
        self.c_ = None
        self.w_ = None
        self.intercept_ = None
        self.coef_ = None
        self.pattern_ = None
        self.metric_ = metrics.r2

    # -- the implicit design ------------------------------------------------------------------------
    def _design(self, Xd: torch.Tensor, trial_idx=None, feat_idx=None) -> _lib.Design:
        """Pad so that lag w of time t is padded sample t + w (edges replicated like the reference's
        clamped gather, timedelayed.py:415-417)."""
        xp = _lib.trf_pad(Xd.contiguous(), -self._lag_min, self._lag_max)
        return _lib.Design(xp, n_time=Xd.shape[2], n_lag=self._lag_max - self._lag_min + 1, trial_idx=trial_idx, feat_idx=feat_idx)

    def _finish_fit(self, fit: Dict, design: _lib.Design, yd: torch.Tensor, out_dev) -> None:
        est = self.estimator
        est.coef_ = fit['coef'].to(out_dev)
        est.alpha_ = fit['alpha'].to(out_dev)
        est.loss_ = None if fit['loss'] is None else fit['loss'].to(out_dev)
        est.intercept_ = fit['intercept'][None, :].to(out_dev) if est.fit_intercept else 0.0
        self.intercept_ = est.intercept_
        self.coef_ = est.coef_.reshape((self.f_, self.c_, self.w_)).flip(-1)
        if self.patterns:
            self.pattern_ = self._patterns(fit, design, yd).to(out_dev)

    def _patterns(self, fit: Dict, design: _lib.Design, yd: torch.Tensor) -> torch.Tensor:
        """Haufe patterns cov(X_t) coef^T inv(cov(y_t)), reshaped (f, c, w) *without* transposing, as the
        reference does (timedelayed.py:469-494).  cov(X_t) is the centred Gram / (n - 1) the fit already
        built; the (f x f) target covariance and its inverse are tiny device-side tensor algebra."""
        n = design.n_rows
        gram = fit['gram']
        if gram is None or not self.estimator.fit_intercept:         # the reference always demeans for the patterns
            gram = _lib.trf_stats(_lib.shift_design(design)[0], yd, True)[0]
        Sx = gram / float(n - 1)
        coef = fit['coef']
        yt = yd.permute(0, 2, 1).reshape(-1, yd.shape[1]) if design.trial_idx is None else \
            yd[design.trial_idx.long()].permute(0, 2, 1).reshape(-1, yd.shape[1])
        if yt.shape[1] > 1:
            Py = torch.linalg.inv(_lib.covariance(yt.contiguous()))   # inverse of the (f x f) target covariance
            pat = _lib.matmul_nt(_lib.matmul_nt(Sx, coef), Py.T.contiguous())
        else:
            pat = _lib.matmul_nt(Sx, coef) * (1.0 / float(yt.shape[0] - 1))
        return pat.reshape((self.f_, self.c_, self.w_)).flip(-1)

    def _shared_stats(self, Xd: torch.Tensor, yd: torch.Tensor):
        """Inside ``Hierarchical`` / ``Shapley`` every predictor set is a feature subset of one array, and its Gram and
        cross-moment for a fold are sub-matrices of that fold's full-predictor statistics (the preceding ``Scaler``
        works per feature, so scaling commutes with the selection).  Those are computed once per fold, kept in the
        driver's ``shared`` dict, and sliced here; returns (design on the full padded array with ``feat_idx``, stats)
        or None when the situation is anything else."""
        from sklearn.pipeline import Pipeline
        from ..preprocessing.scaler import _Scaler_torch
        ctx = _lib.fold_ctx
        par, train, model = ctx.parent, ctx.train, ctx.model
        if par is None or train is None or model is None or os.environ.get('MVB_SHARE_STATS', '1') == '0':
            return None
        X_all, idx = par['X'], par['idx']
        if X_all.ndim != 3 or par['dim'] not in (-2, 1) or not X_all.is_cuda or X_all.dtype != Xd.dtype:
            return None
        if isinstance(model, Pipeline):
            if model[-1] is not self:
                return None
            pre = [step for _, step in model.steps[:-1]]
        elif model is self:
            pre = []
        else:
            return None
        for step in pre:              # only per-feature preprocessing commutes with the feature selection
            if not isinstance(step, _Scaler_torch):
                return None
            dims = step.dims if step.dims is not None else (0,)
            dims = (dims,) if isinstance(dims, int) else tuple(int(d) for d in dims)
            if any(d % 3 == 1 for d in dims):
                return None
        n_tr, W = int(train.numel()), self.window.shape[0]
        if Xd.shape != (n_tr, int(idx.numel()), X_all.shape[2]) or n_tr * X_all.shape[2] <= X_all.shape[1] * W + 1:
            return None
        tkey = ctx.train_key if ctx.train_key is not None else tuple(train.tolist())     # the latter reads the indices back (sync)
        key = (tkey, W, self._lag_min, bool(self.estimator.fit_intercept))
        entry = par['shared'].get(key)
        if entry is None:
            X_tr = X_all[train.to(X_all.device)]
            for step in pre:
                X_tr = step.clone().fit_transform(X_tr)
            full, shift = self._design(X_tr.contiguous()), None
            if self.estimator.fit_intercept:
                full, shift = _lib.shift_design(full)             # centring before multiplying (ridgecv.solve_design)
            entry = (full, _lib.trf_stats(full, yd, self.estimator.fit_intercept), shift)
            with _SHARED_LOCK:                                        # folds of one driver may be fitted from several lane threads
                while len(par['shared']) >= SHARED_FOLD_ENTRIES:      # oldest first (dicts keep insertion order): splitters that
                    par['shared'].pop(next(iter(par['shared'])))      # reshuffle per clone would otherwise grow this without reuse
                par['shared'][key] = entry
        full, (G, XtY, mu, ybar), shift = entry
        Gs, Bs, ms = _lib.gather_stats(G, XtY, mu, idx, W)
        return full.subset(feat_idx=idx), (Gs, Bs, ms, ybar), shift

    def fit(self, X: torch.Tensor, y: torch.Tensor):
        if (len(X.shape) != 3) | (len(y.shape) != 3):
            raise ValueError(f'X and y must have 3 dimensions (trials, channels, time), but got {X.shape} and {y.shape}.')
        if (X.shape[0] != y.shape[0]) | (X.shape[2] != y.shape[2]):
            raise ValueError(f'X and y must have the same number of trials and time points, but got {X.shape} and {y.shape}.')
        self.f_, self.c_, self.w_ = y.shape[1], X.shape[1], self.window.shape[0]
        out_dev = X.device
        Xd = _lib.to_device(X)
        yd = _lib.to_device(y).to(Xd.dtype).contiguous()
        est = self.estimator
        est.alphas = est.alphas.to(X.dtype).to(X.device)              # ridgecv.py:125
        shared = self._shared_stats(Xd, yd)
        design, stats, shift = shared if shared is not None else (self._design(Xd), None, None)
        fit = solve_design(design, yd, _lib.to_device(est.alphas).contiguous(), est.fit_intercept, est.alpha_per_target,
                           stats=stats, want_loss=True, keep_gram=self.patterns, shift=shift)
        self._finish_fit(fit, design, yd, out_dev)
        return self

    def predict(self, X: torch.Tensor, reshape: bool = True) -> torch.Tensor:
        if (self.coef_ is None) or (self.intercept_ is None):
            raise ValueError('Estimator has not been fitted yet.')
        if len(X.shape) != 3:
            raise ValueError(f'X must have 3 dimensions (trials, channels, time), but got {X.shape}.')
        out_dev = X.device
        Xd = _lib.to_device(X)
        est = self.estimator
        coef = _lib.to_device(est.coef_).to(Xd.dtype).contiguous()
        icpt = _lib.to_device(est.intercept_).to(Xd.dtype).reshape(-1).contiguous() if isinstance(est.intercept_, torch.Tensor) else None
        yh = _lib.trf_predict(self._design(Xd), coef, icpt)                        # (n, f, t)
        if not reshape:
            yh = yh.permute(0, 2, 1).reshape(-1, self.f_)
        return yh.to(out_dev)

    def score(self, X: torch.Tensor, y: torch.Tensor, metric: Optional[Union[metrics.Metric, Tuple[metrics.Metric]]] = None
              ) -> Union[torch.Tensor, Dict[str, torch.Tensor]]:
        if metric is None:
            metric = self.metric_
        return metrics.score(self, metric, X, y)

    def clone(self):
        return _TimeDelayed_torch(self.t_min, self.t_max, self.fs, alphas=self.alphas, patterns=self.patterns, **self.kwargs)


class TimeDelayed(sklearn.base.BaseEstimator):
    """Dispatcher (timedelayed.py:715-747): scalar / list / tensor alphas -> CUDA estimator."""

    def __new__(cls, t_min: float, t_max: float, fs: int, alphas: Union[torch.Tensor, np.ndarray, float, int] = torch.tensor([1]),
                patterns: bool = False, **kwargs):
        if isinstance(alphas, (float, int)):
            alphas = torch.tensor([alphas])
        if isinstance(alphas, list):
            alphas = torch.tensor(alphas)
        if isinstance(alphas, torch.Tensor):
            return _TimeDelayed_torch(t_min, t_max, fs, alphas=alphas, patterns=patterns, **kwargs)
        raise ValueError(f'Alphas should be of type torch.Tensor (the B200 build has no numpy backend), but got {type(alphas)}.')
