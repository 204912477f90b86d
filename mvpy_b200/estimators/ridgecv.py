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


def _finish_shift(out: Dict, design: _lib.Design, shift: Optional[torch.Tensor]) -> Dict:
    """Undo the pivot shift of ``_lib.shift_design`` in the quantities that depend on it: column means and intercept
    (coefficients, losses and the centred Gram are shift-invariant)."""
    if shift is None:
        return out
    s_col = _lib.shift_columns(design, shift)
    out['intercept'] = out['intercept'] - out['coef'] @ s_col
    out['mu'] = out['mu'] + s_col
    return out


def solve_design(design: _lib.Design, y: torch.Tensor, alphas: torch.Tensor, fit_intercept: bool, alpha_per_target: bool,
                 stats=None, want_loss: bool = False, keep_gram: bool = True, shift: Optional[torch.Tensor] = None):
    """Shared fit core: y is (N, F, T) on the device of ``design``.  Returns dict with flat coef (F, p),
    intercept (F,), alpha (0-d or (F,)), best, loss, and the centred Gram (for patterns).

    ``stats`` / ``shift``: precomputed statistics of a design that was already pivot-shifted by ``shift`` (predictor-set
    drivers share them per fold).  Otherwise, with an intercept, the design is shifted here: the reference centres the
    design before any product (ridgecv.py:59-94, on a clone); forming G - n mu mu^T from uncentred fp32 products instead
    loses (mean/std)^2 eps, so each feature's mean is taken out of the stimulus first (``mvb_trf_shift``)."""
    wide = design.n_rows <= design.n_cols + int(fit_intercept)
    if stats is None and design.xp.dtype == torch.float32 and wide:
        # No more rows than unknowns: every training row is (nearly) interpolated, 1 - leverage is of order
        # alpha / eigenvalue and fp32 cannot form it (the reference's dual route for n <= p, ridgecv.py:135-190, never
        # does).  Small column counts run the primal route on the fp64 kernels; larger ones the dual (n x n) route, fp64 too.
        d64 = _lib.Design(design.xp.double(), design.n_time, design.n_lag, design.trial_idx, design.feat_idx)
        if design.n_cols <= 2048:
            out = solve_design(d64, y.double(), alphas.double(), fit_intercept, alpha_per_target, None, want_loss, keep_gram)
        else:
            out = solve_design_dual(d64, y.double(), alphas.double(), fit_intercept, alpha_per_target, keep_gram)
        return {k: (v.float() if isinstance(v, torch.Tensor) and v.dtype == torch.float64 else v) for k, v in out.items()}
    if stats is None and wide and design.n_cols > 2048:
        return solve_design_dual(design, y, alphas, fit_intercept, alpha_per_target, keep_gram)
    if stats is None:
        if fit_intercept:
            design, shift = _lib.shift_design(design)
        stats = _lib.trf_stats(design, y, fit_intercept)
    G, XtY, mu, ybar = stats
    Gkeep = G.clone() if keep_gram else None          # the solve may overwrite G; patterns need cov(X) = G / (n - 1)
    coef, icpt, alpha, best, loss = _lib.ridge_solve(design, y, G, XtY, mu, ybar, alphas, fit_intercept, alpha_per_target,
                                                     want_loss=want_loss)
    out = dict(coef=coef, intercept=icpt, alpha=alpha if alpha_per_target else alpha.reshape(()), best=best, loss=loss,
               gram=Gkeep, mu=mu, ybar=ybar)
    return _finish_shift(out, design, shift)


DUAL_MAX_ROWS = 2048


def solve_design_dual(design: _lib.Design, y: torch.Tensor, alphas: torch.Tensor, fit_intercept: bool, alpha_per_target: bool,
                      keep_gram: bool = False):     # the p x p Gram is never built here: callers that need it (patterns) rebuild it
    """n <= p: the reference's dual route (ridgecv.py:135-205) -- eigendecomposition of the n x n kernel matrix
    K = X_c X_c^T (+ 1 1^T for the unpenalised constant direction), duals = Q diag(w) Q^T y_c, LOO error = dual / diag,
    coef = duals^T X_c.  K and duals^T X come from the contraction kernel over the implicit design (both orientations),
    the eigensolve from the batched Jacobi kernel; everything else is n x n sized tensor algebra.  fp64 only (the caller
    promotes fp32 inputs); at most ``DUAL_MAX_ROWS`` rows -- beyond that the n x n eigensolve has no kernel here."""
    if design.xp.dtype != torch.float64:
        raise _lib.MvbError('solve_design_dual works on float64 designs')
    n, p, F = design.n_rows, design.n_cols, y.shape[1]
    if n > DUAL_MAX_ROWS:
        raise _lib.MvbError(f'ridge fit with no more rows than columns (n = {n} <= p = {p}) and more than {DUAL_MAX_ROWS} rows: the dual '
                            f'(n x n) eigensolve this regime needs (ridgecv.py:135-190) is only implemented up to {DUAL_MAX_ROWS} rows, '
                            f'and the Gram route cannot form 1 - leverage there.  Use more training samples than columns.')
    dev, dt = design.xp.device, design.xp.dtype
    shift = None
    if fit_intercept:
        design, shift = _lib.shift_design(design)
    xp = design.xp
    tri = torch.arange(design.n_trials, device=dev) if design.trial_idx is None else design.trial_idx.long()
    fti = torch.arange(design.n_feat, device=dev) if design.feat_idx is None else design.feat_idx.long()
    row_off = (tri[:, None] * xp.stride(0) + torch.arange(design.n_time, device=dev)[None, :]).reshape(-1).contiguous()
    col_off = (fti[:, None] * xp.stride(1) + torch.arange(design.n_lag, device=dev)[None, :]).reshape(-1).contiguous()
    yt = (y if design.trial_idx is None else y[tri]).permute(0, 2, 1).reshape(n, F)
    K = _lib.contract(xp, row_off, col_off, 0, xp, row_off, col_off, 0, n, n, p, symmetric=True)         # X X^T
    ones_off = torch.zeros(1, dtype=torch.int64, device=dev)
    ones = torch.ones(n, dtype=dt, device=dev)
    k_iota = torch.arange(n, device=dev)
    col_sum = _lib.contract(ones, ones_off, k_iota, 1, xp, col_off, row_off, 0, 1, p, n)[0]              # 1^T X
    if fit_intercept:
        mu, ybar = col_sum / n, yt.mean(0)
        v = K.mean(1)                                       # X mu
        K = K - v[:, None] - v[None, :] + v.mean() + 1.0    # X_c X_c^T + 1 1^T   (ridgecv.py:139-141)
        yc = yt - ybar
    else:
        mu, ybar = torch.zeros(p, dtype=dt, device=dev), torch.zeros(F, dtype=dt, device=dev)
        yc = yt
    Vt, lam = _lib.eigh(K[None].contiguous())
    Q, lam = Vt[0].mT, lam[0]
    Qty = Q.mT @ yc
    W = 1.0 / (lam[None, :] + alphas[:, None])              # (A, n)
    if fit_intercept:
        unit = torch.full((n,), n ** -0.5, dtype=dt, device=dev)
        k0 = int((Q.mT @ unit).abs().argmax())              # the constant direction is not penalised (ridgecv.py:147-150)
        W[:, k0] = 0.0
    duals = torch.einsum('ik,ak,kf->aif', Q, W, Qty)
    diag = (Q * Q) @ W.mT                                   # (n, A)
    loss = (duals / diag.mT[:, :, None]).square().mean(1)   # (A, F)
    if alpha_per_target:
        best = loss.argmin(0)
        dual = torch.stack([duals[int(best[f]), :, f] for f in range(F)], 1)
        alpha = alphas[best]
    else:
        best = loss.mean(1).argmin().reshape(1)
        dual = duals[int(best)]
        alpha = alphas[best].reshape(())
    dT = dual.mT.contiguous()                               # (F, n)
    f_off = torch.arange(F, device=dev) * n
    coef = _lib.contract(dT, f_off, k_iota, 1, xp, col_off, row_off, 0, F, p, n)                        # duals^T X
    if fit_intercept:
        coef = coef - dT.sum(1)[:, None] * mu[None, :]      # ... - (duals^T 1) mu^T = duals^T X_c
    icpt = ybar - coef @ mu if fit_intercept else torch.zeros(F, dtype=dt, device=dev)
    out = dict(coef=coef.contiguous(), intercept=icpt, alpha=alpha, best=best.to(torch.int32), loss=loss, gram=None, mu=mu, ybar=ybar)
    return _finish_shift(out, design, shift)


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
