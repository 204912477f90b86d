"""Receptive-field estimation (mTRF / stimulus reconstruction from per-epoch correlations) -- drop-in for
``mvpy.estimators.ReceptiveField`` (mvpy/estimators/receptivefield.py:832-1837), SURVEY.md section 8(f)-1.

Model:  y[n, f, t] = sum_c sum_k coef_[f, c, k] * X[n, c, t - (s_min + k)] + intercept_[f]   (X zero outside the epoch).

The reference forms, per epoch, the lag auto-correlation matrix and the cross-correlation with y by FFT and removes
the rows outside the epoch again ("edge correction").  Those matrices are exactly X_lag^T X_lag and X_lag^T y of the
zero-padded lag design restricted to the epoch's samples, so here they come from the implicit-design contraction kernel
(``mvb_trf_stats``, tcgen05 split-TF32 for fp32 / CUDA cores for fp64) on a zero-padded copy of X, one epoch at a
time; ``mvb_rf_assemble`` then lays them out as the reference's ``cov_`` (and reproduces its two deviations from the
textbook matrices, see csrc/rf_kernels.cuh), ``mvb_sum_slabs`` sums the epochs of a fold, ``mvb_penalised_solve``
solves the penalised normal equations for the whole alpha grid, ``mvb_trf_predict`` / ``mvb_score`` evaluate the
held-out epochs.  Nothing is computed on the host."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import sklearn.base
import torch

from .. import _lib, metrics
from ..crossvalidation import KFold
from ..preprocessing.scaler import _Scaler_torch
from .ridgecv import _RidgeCV_torch


def _free_end_laplacian(m: int) -> torch.Tensor:
    """Second-difference operator of a chain of m nodes with free ends: diag (1, 2, ..., 2, 1), off-diagonals -1."""
    deg = torch.full((m,), 2.0)
    deg[0] = deg[-1] = 1.0
    link = torch.ones(m - 1)
    return torch.diag(deg) - torch.diag(link, 1) - torch.diag(link, -1)


class _ReceptiveField_torch(sklearn.base.BaseEstimator):
    """Attributes: ``t_min, t_max, fs, alpha, reg_type, reg_cv, patterns, fit_intercept, edge_correction, s_min, s_max,
    window``; after ``fit``: ``n_features_, n_channels_, n_trf_, cov_`` (n, c*w, c*w), ``coef_`` (f, c, w),
    ``intercept_`` (f,) or 0.0, ``pattern_``, ``alpha_``; ``metric_`` = r2."""

    def __init__(self, t_min: float, t_max: float, fs: int, alpha: Union[float, torch.Tensor, List] = 1.0,
                 reg_type: Union[str, List] = 'ridge', reg_cv: Any = 5, patterns: bool = False, fit_intercept: bool = True,
                 edge_correction: bool = True):
        self.t_min, self.t_max, self.fs = t_min, t_max, fs
        self.alpha = alpha
        self.reg_type = reg_type
        self.reg_cv = reg_cv
        self.patterns = patterns
        self.fit_intercept = fit_intercept
        self.edge_correction = edge_correction
        if t_max < t_min:
            raise ValueError(f't_min must be smaller than t_max, but got {t_min} > {t_max}.')
        if (reg_type != 'ridge') and (reg_cv == 'LOO'):
            raise ValueError(f'reg_cv=`LOO` is available only for reg_type=`ridge`, but got {reg_type}.')
        if not hasattr(reg_cv, 'split') and not isinstance(reg_cv, (int, str)):
            raise ValueError('`reg_cv` must either be of type int specifying number of folds to use, '
                             'str (\'LOO\') for RidgeCV solving, or expose a `split` method for cross-validation.')
        if isinstance(self.alpha, list):
            self.alpha = torch.tensor(self.alpha)
        if isinstance(self.alpha, torch.Tensor) and self.alpha.shape[0] == 1:
            self.alpha = float(self.alpha.cpu().item())
        if not isinstance(self.alpha, (torch.Tensor, float)):
            raise ValueError(f'Unknown `alpha` type, must be either float, List, or torch.Tensor, but got {type(self.alpha)}.')
        self.s_min = int(t_min * fs)
        self.s_max = int(t_max * fs) + 1
        self.window = torch.arange(self.s_min, self.s_max)
        self.n_features_ = None
        self.n_channels_ = None
        self.n_trf_ = None
        self.cov_ = None
        self.coef_ = None
        self.intercept_ = None
        self.pattern_ = None
        self.alpha_ = None
        self.metric_ = metrics.r2

    # -- the zero-padded lag design -----------------------------------------------------------------
    @staticmethod
    def _shift(x: torch.Tensor, left: int, length: int) -> torch.Tensor:
        """out[..., j] = x[..., j - left] for j in [0, length), zero where x has no sample."""
        T = x.shape[-1]
        out = x.new_zeros(x.shape[:-1] + (length,))
        lo, hi = max(left, 0), min(left + T, length)
        if hi > lo:
            out[..., lo:hi] = x[..., lo - left:hi - left]
        return out

    def _design(self, Xd: torch.Tensor, extended: bool = False, trial_idx=None) -> _lib.Design:
        """Kernel lag index w reads padded sample t' + w; filter lag k = W-1-w.  Rows are the epoch's own samples, or
        (extended) every row that any lagged sample reaches -- the full correlation sums without edge correction."""
        W, T = self.s_max - self.s_min, Xd.shape[-1]
        rows = T + W - 1 if extended else T
        left = W - 1 if extended else self.s_max - 1
        xp = self._shift(Xd, left, rows + W - 1).contiguous()
        return _lib.Design(xp, n_time=rows, n_lag=W, trial_idx=trial_idx)

    def _compute_corrs(self, Xd: torch.Tensor, yd: torch.Tensor):
        """cov (n, p, p) in the reference layout, xty (n, p, f) with row c*W+k, offsets (c,), (f,)
        (receptivefield.py:976-1118)."""
        N, C, T = Xd.shape
        F = yd.shape[1]
        W = self.s_max - self.s_min
        self.n_features_, self.n_channels_, self.n_trf_ = C, F, W
        if self.fit_intercept:
            sx = _Scaler_torch(with_mean=True, with_std=False, dims=(0, 2)).fit(Xd)
            sy = _Scaler_torch(with_mean=True, with_std=False, dims=(0, 2)).fit(yd)
            x_off, y_off = sx.mean_.reshape(-1), sy.mean_.reshape(-1)
            Xd, yd = sx.transform(Xd).contiguous(), sy.transform(yd).contiguous()
        else:
            x_off = y_off = None
        extended = not self.edge_correction
        design = self._design(Xd, extended)
        y_rows = self._shift(yd, -self.s_min, design.n_time).contiguous() if extended else yd
        p = C * W
        gram = torch.empty((N, p, p), dtype=Xd.dtype, device=Xd.device)
        xty = torch.empty((N, p, F), dtype=Xd.dtype, device=Xd.device)
        epochs = torch.arange(N, dtype=torch.int32, device=Xd.device)
        for n in range(N):
            one = design.subset(trial_idx=epochs[n:n + 1])
            gram[n], xty[n] = _lib.trf_stats(one, y_rows, False)[:2]
        cov = _lib.rf_assemble(gram, Xd.contiguous(), self.s_min, self.s_max, self.edge_correction)
        xty = xty.reshape(N, C, W, F).flip(2).reshape(N, p, F).contiguous()
        return cov, xty, x_off, y_off

    def _compute_reg_neighbours(self, n_features: int, n_trf: int, reg_type: Union[str, List]) -> torch.Tensor:
        """Penalty matrix (receptivefield.py:1120-1183): identity, or free-end Laplacians over lags (within a feature)
        and / or over features (within a lag); the feature-only case carries no identity."""
        if isinstance(reg_type, str):
            reg_type = (reg_type,) * 2
        if len(reg_type) != 2:
            raise ValueError(f'reg_type must be of length two, but got {len(reg_type)}.')
        for r in reg_type:
            if r not in ['ridge', 'laplacian']:
                raise ValueError(f'Unknown reg_type. Expected ridge or laplacian, but got {r}.')
        over_time = (reg_type[0] == 'laplacian') and (n_trf > 1)
        over_feat = (reg_type[1] == 'laplacian') and (n_features > 1)
        size = n_features * n_trf
        if not over_time and not over_feat:
            return torch.eye(size)
        reg = torch.zeros((size, size))
        if over_time:
            reg = reg + torch.kron(torch.eye(n_features), _free_end_laplacian(n_trf))
        if over_feat:
            reg = reg + torch.kron(_free_end_laplacian(n_features), torch.eye(n_trf))
        return reg

    def _solve(self, G: torch.Tensor, b: torch.Tensor, alphas: torch.Tensor, reg: torch.Tensor) -> torch.Tensor:
        """(G + alpha reg) w = b for every alpha -> (S, f, c, w) (receptivefield.py:1223-1244)."""
        sol, info = _lib.penalised_solve(G, reg, alphas, b)
        bad = torch.nonzero(info).reshape(-1).tolist()
        for s in bad:            # exactly singular system: minimum-norm least squares, as the reference falls back to (:1230-1238)
            sol[s] = torch.linalg.lstsq(G + alphas[s] * reg, b).solution
        return sol.transpose(1, 2).reshape(alphas.numel(), self.n_channels_, self.n_features_, self.n_trf_)

    def _fit_intercept(self, x_off, y_off, coef: torch.Tensor):
        if not self.fit_intercept:
            return 0.0
        return y_off - x_off @ coef.sum(-1).T

    def _predict(self, Xd: torch.Tensor, coef: torch.Tensor, intercept, design: Optional[_lib.Design] = None) -> torch.Tensor:
        """(n, f, t) predictions of the lagged model on the device (receptivefield.py:1434-1488)."""
        if design is None:
            design = self._design(Xd)
        Fn = coef.shape[0]
        flat = coef.flip(-1).reshape(Fn, -1).to(design.xp.dtype).contiguous()           # kernel lag order
        icpt = intercept.to(design.xp.dtype).reshape(-1).contiguous() if isinstance(intercept, torch.Tensor) else None
        return _lib.trf_predict(design, flat, icpt)

    def _fit_pattern(self, cov: torch.Tensor, yd: torch.Tensor, coef: torch.Tensor) -> Optional[torch.Tensor]:
        """cov(X_lag) coef pinv(cov(y)), with the reference's plain reshape of coef (receptivefield.py:1277-1320)."""
        if not self.patterns:
            return None
        N, F, T = yd.shape
        p = cov.shape[1]
        S_X = _lib.sum_slabs(cov, torch.arange(N)) / float(T * N - 1)
        if F > 1:
            z = yd.swapaxes(1, 2).reshape(-1, F).contiguous()
            P_y = torch.linalg.pinv(_lib.covariance(z))                 # (f x f) target covariance: tiny
            M = _lib.matmul_nt(coef.reshape(p, F).contiguous(), P_y.T.contiguous())
        else:
            M = coef.reshape(p, F) * (1.0 / float(T * N - 1))
        return _lib.matmul_nt(S_X, M.T.contiguous()).reshape(coef.shape)

    # -- public API -------------------------------------------------------------------------------
    def fit(self, X: torch.Tensor, y: torch.Tensor):
        if len(X.shape) not in [2, 3]:
            raise ValueError(f'`X` must be of shape (n_samples[, n_features], n_timepoints), but got {X.shape}.')
        if len(X.shape) == 2:
            X = X[:, None, :]
        if len(y.shape) not in [2, 3]:
            raise ValueError(f'`y` must be of shape (n_samples[, n_channels], n_timepoints), but got {y.shape}.')
        if len(y.shape) == 2:
            y = y[:, None, :]
        if (X.shape[0] != y.shape[0]) or (X.shape[-1] != y.shape[-1]):
            raise ValueError('`X` (n_samples[, n_features], n_timepoints) and `y` (n_samples[, n_channels], n_timepoints) '
                             f'must match in (n_samples, n_timepoints), but got X={(X.shape[0], X.shape[-1])} and '
                             f'y={(y.shape[0], y.shape[-1])}.')
        out_dev = X.device
        Xd = _lib.to_device(X).contiguous()
        yd = _lib.to_device(y).to(Xd.dtype).contiguous()
        N = Xd.shape[0]
        cov, xty, x_off, y_off = self._compute_corrs(Xd, yd)
        self.cov_ = cov.to(out_dev)
        everything = torch.arange(N)

        def finish(coef: torch.Tensor, alpha_) -> '_ReceptiveField_torch':
            icpt = self._fit_intercept(x_off, y_off, coef)
            self.coef_ = coef.to(out_dev)
            self.alpha_ = alpha_
            self.intercept_ = icpt.to(out_dev) if isinstance(icpt, torch.Tensor) else icpt
            pat = self._fit_pattern(cov, yd, coef)
            self.pattern_ = None if pat is None else pat.to(out_dev)
            return self

        G_all, b_all = _lib.sum_slabs(cov, everything), _lib.sum_slabs(xty, everything)
        if self.reg_cv == 'LOO':
            # The summed correlation matrix is handed to the LOO ridge solver as its data matrix (:1213-1221).  That
            # "design" is square: the ridge core switches to its fp64 kernels for it (estimators/ridgecv.py, solve_design).
            est = _RidgeCV_torch(alphas=self.alpha, fit_intercept=True).fit(G_all, b_all)
            coef = est.coef_.reshape((self.n_channels_, self.n_features_, self.n_trf_))
            return finish(coef, est.alpha_.to(out_dev))
        reg = self._compute_reg_neighbours(self.n_features_, self.n_trf_, self.reg_type).to(device=Xd.device, dtype=Xd.dtype)
        if not isinstance(self.alpha, float):
            if self.alpha.device != X.device:
                self.alpha = self.alpha.to(X.device)
            alphas = _lib.to_device(self.alpha).to(Xd.dtype).contiguous()
            cv = self.reg_cv if hasattr(self.reg_cv, 'split') else KFold(n_splits=self.reg_cv).to_torch()
            folds = list(cv.split(cov, xty))
            design = self._design(Xd)
            scores = torch.empty((alphas.numel(), len(folds)), dtype=Xd.dtype, device=Xd.device)
            for f_i, (train, test) in enumerate(folds):
                coefs = self._solve(_lib.sum_slabs(cov, train), _lib.sum_slabs(xty, train), alphas, reg)
                held = design.subset(trial_idx=test)
                y_test = yd[test.to(yd.device)].reshape(test.numel(), -1).contiguous()
                for a_i in range(alphas.numel()):
                    y_h = self._predict(Xd, coefs[a_i], self._fit_intercept(x_off, y_off, coefs[a_i]), held)
                    # correlation across the held-out epochs for every (channel, time) cell, then the mean (:1393-1397)
                    scores[a_i, f_i] = _lib.score(y_test, y_h.reshape(test.numel(), -1), None, False, True)[1].mean()
            per_fold = scores.cpu().tolist()
            oos = torch.tensor([torch.mean(torch.tensor(row)).item() for row in per_fold])
            best = torch.argmax(oos)
            self.alpha_ = self.alpha[best]
            chosen = alphas[best:best + 1]
        else:
            self.alpha_ = self.alpha
            chosen = torch.tensor([self.alpha], dtype=Xd.dtype, device=Xd.device)
        return finish(self._solve(G_all, b_all, chosen, reg)[0], self.alpha_)

    def predict(self, X: torch.Tensor) -> torch.Tensor:
        if self.coef_ is None or self.intercept_ is None:
            raise ValueError('The estimator has not been fitted yet.')
        if len(X.shape) == 2:
            X = X[:, None, :]
        Xd = _lib.to_device(X).contiguous()
        icpt = _lib.to_device(self.intercept_) if isinstance(self.intercept_, torch.Tensor) else self.intercept_
        y_h = self._predict(Xd, _lib.to_device(self.coef_), icpt)
        return y_h.to(torch.float32).to(X.device)               # the reference accumulates into a float32 buffer (:1460)

    def score(self, X: torch.Tensor, y: torch.Tensor, metric: Optional[Union[metrics.Metric, Tuple[metrics.Metric]]] = None
              ) -> Union[torch.Tensor, Dict[str, torch.Tensor]]:
        if metric is None:
            metric = self.metric_
        return metrics.score(self, metric, X, y)

    def clone(self):
        return _ReceptiveField_torch(self.t_min, self.t_max, self.fs, alpha=self.alpha, reg_type=self.reg_type, reg_cv=self.reg_cv,
                                     patterns=self.patterns, fit_intercept=self.fit_intercept, edge_correction=self.edge_correction)


class ReceptiveField(sklearn.base.BaseEstimator):
    """Dispatcher (receptivefield.py:1718-1767): int / float / list / tensor alpha -> CUDA estimator."""

    def __new__(cls, t_min: float, t_max: float, fs: int, alpha: Union[int, float, np.ndarray, torch.Tensor, List] = 1.0,
                reg_type: Union[str, List] = 'ridge', reg_cv: Any = 5, patterns: bool = False, fit_intercept: bool = True,
                edge_correction: bool = True):
        if isinstance(alpha, (float, int)):
            alpha = torch.tensor([float(alpha)])
        if isinstance(alpha, list):
            alpha = torch.tensor(alpha)
        if isinstance(alpha, torch.Tensor):
            return _ReceptiveField_torch(t_min, t_max, fs, alpha=alpha, reg_type=reg_type, reg_cv=reg_cv, patterns=patterns,
                                         fit_intercept=fit_intercept, edge_correction=edge_correction)
        raise ValueError('Unknown type of alpha. Expected int, float, list or torch.Tensor (the B200 build has no numpy '
                         f'backend), but got {type(alpha)}.')
