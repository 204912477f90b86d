"""CPU oracle for the cross-validated ridge hot path of FabulousFabs/MVPy.

TEST INFRASTRUCTURE ONLY.  This module restates, on CPU torch, the algorithms the reference
uses for the path named in BASELINE.json (north_star).  It is imported only by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``;
nothing under ``mvpy_b200/`` may import it, and the product never falls back to it.

Pinning: ``tests/test_oracle_golden.py`` checks every function here against fixtures produced by
the unmodified reference (``oracle/gen_golden.py`` imports ``/root/reference`` and writes
``tests/golden/*.npz``), including the reference's own known-answer cases
(``tests/estimators/test_ridgecv.py`` sizes, the ``_delay_matrices`` arange example, the
docstring KFold listing).  Every function cites the reference lines it follows.

Two solvers are provided for the ridge core:
  * ``ridge_loo_fit``        -- the reference's own route (dual eigh for n<=p, thin SVD of [X,1]
                                otherwise; ``mvpy/estimators/ridgecv.py:96-303``).  Used as the CPU
                                baseline because it is what the reference spends its time in.
  * ``ridge_loo_fit_primal`` -- the algebraically equivalent primal route (eigendecomposition of the
                                centred Gram + per-row leverages, SURVEY.md section 8 a-note) that the
                                CUDA kernels implement; fp64 spec for kernel tests.
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------------------------
# lag window and the lag-expanded design            (mvpy/estimators/timedelayed.py:376-431)
# --------------------------------------------------------------------------------------------


def lag_window(t_min: float, t_max: float, fs: float) -> np.ndarray:
    """Integer lags spanned by (t_min, t_max) at sampling rate fs.

    reference: timedelayed.py:376-380 -- both bounds go through abs(ceil(.)), the negative
    side is mirrored, so the window always contains 0 and runs -|ceil(t_min fs)| .. +|ceil(t_max fs)|.
    """
    n_neg = int(abs(math.ceil(t_min * fs)))
    n_pos = int(abs(math.ceil(t_max * fs)))
    return np.arange(-n_neg, n_pos + 1, dtype=np.int64)


def delay_design(X: torch.Tensor, window: np.ndarray) -> torch.Tensor:
    """(N, C, T) -> (N*T, C*W) with column c*W+w = X[n, c, clamp(t + window[w], 0, T-1)].

    reference: timedelayed.py:394-423.  Out-of-range lags are *clamped* (edge samples are
    replicated); the zeroing mask at :420 can never fire because it is applied after the clamp.
    """
    N, C, T = X.shape
    w = torch.as_tensor(window, dtype=torch.long, device=X.device)
    t_idx = (torch.arange(T, device=X.device)[None, :] + w[:, None]).clamp_(0, T - 1)  # (W, T)
    gathered = X[:, :, t_idx]  # (N, C, W, T)
    return gathered.permute(0, 3, 1, 2).reshape(N * T, C * w.numel())


def flatten_targets(y: torch.Tensor) -> torch.Tensor:
    """(N, F, T) -> (N*T, F)                      reference: timedelayed.py:426-428."""
    return y.permute(0, 2, 1).reshape(-1, y.shape[1])


# --------------------------------------------------------------------------------------------
# LOO ridge: the reference's route                    (mvpy/estimators/ridgecv.py:59-303)
# --------------------------------------------------------------------------------------------


def _select(losses: torch.Tensor, alphas: torch.Tensor, per_target: bool):
    """first-minimum rule of ridgecv.py:174-177,190-192 / 253-256,271-274."""
    if per_target:
        best = losses.argmin(0)
    else:
        best = losses.mean(1).argmin()
    return best, alphas[best]


def ridge_loo_fit(X: torch.Tensor, y: torch.Tensor, alphas: torch.Tensor, fit_intercept: bool = True,
                  alpha_per_target: bool = False) -> Dict[str, torch.Tensor]:
    """Exact leave-one-out ridge over an alpha grid; returns coef (f,p), intercept, alpha, losses (A,f).

    reference: ridgecv.py:96-303 (``_RidgeCV_torch.fit``): centre (:81-88), dual Gram + eigh when
    n <= p (:135-205), thin SVD of [X, 1] otherwise (:206-290), LOO error = dual / diag(G^-1)
    (:171, :250), mean over samples (:172, :251), coefficients = duals^T X (:205, :290),
    intercept = y_mean - X_mean coef^T (:297-299).
    """
    if y.ndim == 1:
        y = y[:, None]
    n, p = X.shape
    dt = X.dtype
    alphas = torch.as_tensor(alphas).to(dt).reshape(-1)
    if fit_intercept:
        x_mu, y_mu = X.mean(0), y.mean(0)
        Xc, yc = X - x_mu, y - y_mu
    else:
        x_mu, y_mu = torch.zeros(p, dtype=dt), torch.zeros(y.shape[1], dtype=dt)
        Xc, yc = X.clone(), y.clone()
    one = torch.ones(n, dtype=dt)
    unit = one / one.norm()

    if n <= p:
        K = Xc @ Xc.mT
        if fit_intercept:
            K = K + 1.0  # the unregularised constant direction (:139-141)
        lam, Q = torch.linalg.eigh(K)
        Qty = Q.mT @ yc
        k0 = (Q.mT @ unit).abs().argmax() if fit_intercept else None

        def inv_for(a):  # a: (m,) -> (m, n)
            w = 1.0 / (lam[None, :] + a[:, None])
            if k0 is not None:
                w[:, k0] = 0.0
            return w

        Wg = inv_for(alphas)  # (A, n)
        duals = torch.einsum('ik,ak,kf->aif', Q, Wg, Qty)
        diag = (Q * Q) @ Wg.mT  # (n, A)
        losses = (duals / diag.mT[:, :, None]).square().mean(1)
        best, chosen = _select(losses, alphas, alpha_per_target)
        Wc = inv_for(chosen.reshape(-1))  # (1 or f, n)
        if alpha_per_target:
            dual = Q @ (Wc.mT * Qty)
        else:
            dual = Q @ (Wc[0][:, None] * Qty)
    else:
        D = torch.cat((Xc, one[:, None]), 1) if fit_intercept else Xc
        U, s, _ = torch.linalg.svd(D, full_matrices=False)
        s2 = s * s
        Uty = U.mT @ yc
        k0 = (U.mT @ unit).abs().argmax() if fit_intercept else None

        def w_for(a):  # (m,) -> (m, r)
            w = 1.0 / (s2[None, :] + a[:, None]) - 1.0 / a[:, None]
            if k0 is not None:
                w[:, k0] = -1.0 / a
            return w

        Wg = w_for(alphas)
        ainv = 1.0 / alphas
        duals = torch.einsum('ik,ak,kf->aif', U, Wg, Uty) + ainv[:, None, None] * yc[None]
        diag = (U * U) @ Wg.mT + ainv[None, :]  # (n, A)
        losses = (duals / diag.mT[:, :, None]).square().mean(1)
        best, chosen = _select(losses, alphas, alpha_per_target)
        Wc = w_for(chosen.reshape(-1))
        if alpha_per_target:
            dual = U @ (Wc.mT * Uty) + yc / chosen[None, :]
        else:
            dual = U @ (Wc[0][:, None] * Uty) + yc / chosen
    coef = dual.mT @ Xc
    intercept = (y_mu - x_mu[None, :] @ coef.mT) if fit_intercept else 0.0
    return dict(coef=coef, intercept=intercept, alpha=chosen, best=best, losses=losses)


def ridge_loo_fit_primal(X: torch.Tensor, y: torch.Tensor, alphas: torch.Tensor, fit_intercept: bool = True,
                         alpha_per_target: bool = False) -> Dict[str, torch.Tensor]:
    """Same quantities from the centred Gram G = Xc^T Xc = V L V^T (the route the kernels take).

    r_a = yc - Z (w_a o V^T b),  w_a = 1/(L + a),  Z = Xc V,  b = Xc^T yc,
    d_a,i = 1 - sum_k Z_ik^2 w_a,k - [fit_intercept]/n,  loss[a,f] = mean_i (r_a,if / d_a,i)^2.
    Equivalent to ``ridge_loo_fit`` in both of its branches (SURVEY.md section 8 a-note).
    """
    if y.ndim == 1:
        y = y[:, None]
    n, p = X.shape
    dt = X.dtype
    alphas = torch.as_tensor(alphas).to(dt).reshape(-1)
    if fit_intercept:
        x_mu, y_mu = X.mean(0), y.mean(0)
    else:
        x_mu, y_mu = torch.zeros(p, dtype=dt), torch.zeros(y.shape[1], dtype=dt)
    Xc, yc = X - x_mu, y - y_mu
    G = Xc.mT @ Xc
    b = Xc.mT @ yc
    lam, V = torch.linalg.eigh(G)
    lam = lam.clamp_min(0)
    Z = Xc @ V
    Vtb = V.mT @ b
    c0 = 1.0 / n if fit_intercept else 0.0
    W = 1.0 / (lam[None, :] + alphas[:, None])  # (A, p)
    lev = (Z * Z) @ W.mT  # (n, A)
    d = 1.0 - lev - c0
    losses = torch.stack([((yc - Z @ (W[a][:, None] * Vtb)) / d[:, a:a + 1]).square().mean(0)
                          for a in range(alphas.numel())])
    best, chosen = _select(losses, alphas, alpha_per_target)
    if alpha_per_target:
        coef = (V @ (Vtb / (lam[:, None] + chosen[None, :]))).mT
    else:
        coef = (V @ (Vtb / (lam[:, None] + chosen))).mT
    intercept = (y_mu - x_mu[None, :] @ coef.mT) if fit_intercept else 0.0
    return dict(coef=coef, intercept=intercept, alpha=chosen, best=best, losses=losses, lam=lam, V=V, G=G, b=b)


def ridge_loo_fit_chol(X: torch.Tensor, y: torch.Tensor, alphas: torch.Tensor, fit_intercept: bool = True,
                       alpha_per_target: bool = False) -> Dict[str, torch.Tensor]:
    """The same LOO quantities with one Cholesky factorisation per alpha instead of an eigendecomposition:
    beta_a = (G + a I)^-1 b,  h_a,i = x_i^T (G + a I)^-1 x_i + [fit_intercept]/n = |L_a^-1 x_i|^2 + c0,
    loss[a, f] = mean_i ((y_c - X_c beta_a)_if / (1 - h_a,i))^2           (ridgecv.py:223-256 in primal form).
    Device-agnostic (runs in fp64 on the GPU in seconds at BASELINE configs[1] size, where a 12 864-wide fp64
    eigendecomposition takes a minute) and pinned against ``ridge_loo_fit`` / the golden fixtures in
    tests/test_oracle_golden.py."""
    if y.ndim == 1:
        y = y[:, None]
    n, p = X.shape
    dt, dev = X.dtype, X.device
    alphas = torch.as_tensor(alphas).to(dt).to(dev).reshape(-1)
    if fit_intercept:
        x_mu, y_mu = X.mean(0), y.mean(0)
    else:
        x_mu, y_mu = torch.zeros(p, dtype=dt, device=dev), torch.zeros(y.shape[1], dtype=dt, device=dev)
    Xc, yc = X - x_mu, y - y_mu
    G = Xc.mT @ Xc
    b = Xc.mT @ yc
    XcT = Xc.mT.contiguous()
    c0 = 1.0 / n if fit_intercept else 0.0
    eye = torch.eye(p, dtype=dt, device=dev)
    losses, betas = [], []
    for a in alphas.tolist():
        L = torch.linalg.cholesky(G + a * eye)
        beta = torch.cholesky_solve(b, L)                                  # (p, f)
        S = torch.linalg.solve_triangular(L, XcT, upper=False)            # (p, n) = L^-1 X_c^T
        h = S.square().sum(0) + c0
        del S
        losses.append(((yc - Xc @ beta) / (1.0 - h)[:, None]).square().mean(0))
        betas.append(beta)
    losses = torch.stack(losses)
    best, chosen = _select(losses, alphas, alpha_per_target)
    if alpha_per_target:
        coef = torch.stack([betas[int(best[f])][:, f] for f in range(y.shape[1])])
    else:
        coef = betas[int(best)].mT.contiguous()
    intercept = (y_mu - x_mu[None, :] @ coef.mT) if fit_intercept else 0.0
    return dict(coef=coef, intercept=intercept, alpha=chosen, best=best, losses=losses, G=G, b=b)


def ridge_predict(X: torch.Tensor, coef: torch.Tensor, intercept) -> torch.Tensor:
    """reference: ridgecv.py:323."""
    return X @ coef.mT + intercept


# --------------------------------------------------------------------------------------------
# Scaler                                                (mvpy/preprocessing/scaler.py:301-395)
# --------------------------------------------------------------------------------------------


def scaler_fit(X: torch.Tensor, dims: Sequence[int] = (0,)) -> Dict[str, torch.Tensor]:
    """mean / unbiased variance over ``dims`` (keepdim), scale = sqrt(var), zero or NaN scale -> 1.

    reference: scaler.py:355-360 (no sample_weight branch).
    """
    dims = tuple(int(d) for d in dims)
    mean = X.mean(dim=dims, keepdim=True)
    var = X.var(dim=dims, keepdim=True)
    scale = var.sqrt()
    scale[(scale == 0) | scale.isnan()] = 1.0
    return dict(mean=mean, var=var, scale=scale)


def scaler_transform(X: torch.Tensor, st: Dict[str, torch.Tensor], with_mean=True, with_std=True) -> torch.Tensor:
    """reference: scaler.py:385-393."""
    out = X.clone()
    if with_mean:
        out = out - st['mean']
    if with_std:
        out = out / st['scale']
    return out


# --------------------------------------------------------------------------------------------
# K-fold indices                                        (mvpy/crossvalidation/kfold.py:189-249)
# --------------------------------------------------------------------------------------------


def kfold_indices(n: int, k: int, shuffle: bool = False, seed: Optional[int] = None) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Contiguous test blocks of size n//k (+1 for the first n%k folds) over arange(n) or a
    seeded ``torch.randperm`` (CPU generator).  reference: kfold.py:205-249."""
    if n < k:
        raise ValueError('n_samples must be at least n_splits')
    if shuffle:
        g = torch.Generator()
        if seed is not None:
            g.manual_seed(seed)
        order = torch.randperm(n, generator=g).long()
    else:
        order = torch.arange(n).long()
    sizes = [n // k + (1 if i < n % k else 0) for i in range(k)]
    out, lo = [], 0
    for sz in sizes:
        hi = lo + sz
        test = order[lo:hi]
        train = torch.cat((order[:lo], order[hi:]))
        out.append((train, test))
        lo = hi
    return out


# --------------------------------------------------------------------------------------------
# metrics                             (mvpy/math/r2.py:57-101, mvpy/math/pearsonr.py:35-56,
#                                      mvpy/metrics/score.py:17-46)
# --------------------------------------------------------------------------------------------


def move_reduce_last(x: torch.Tensor, dims: Sequence[int]) -> torch.Tensor:
    """metrics/score.py:17-46 (``reduce_``): the reduce dims become one flattened trailing dim."""
    dims = sorted(d % x.ndim for d in dims)
    keep = [i for i in range(x.ndim) if i not in dims]
    xp = x.permute(*keep, *dims)
    return xp.reshape(*[x.shape[i] for i in keep], -1)


def r2_last(y: torch.Tensor, y_h: torch.Tensor, y_b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """1 - SS_res/SS_tot over the last dim; baseline y_b (default: mean of y); 0 where SS_tot <= 0.
    reference: math/r2.py:80-100."""
    if y_b is None:
        y_b = y.mean(-1, keepdim=True)
    num = (y - y_h).square().sum(-1)
    den = (y - y_b).square().sum(-1)
    out = torch.zeros_like(num)
    ok = den > 0
    out[ok] = 1 - num[ok] / den[ok]
    return out


def pearsonr_last(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """reference: math/pearsonr.py:55-56 (no zero-variance guard: NaN there)."""
    dx, dy = x - x.mean(-1, keepdim=True), y - y.mean(-1, keepdim=True)
    return (dx * dy).sum(-1) / torch.sqrt(dx.square().sum(-1) * dy.square().sum(-1))


def rank_last(x: torch.Tensor) -> torch.Tensor:
    """Average ranks (1-based, ties share the mean of their positions) along the last dim.
    reference: math/rank.py:62-112 (sort, first / last position of each run of equal values, mean of the two)."""
    if not torch.is_floating_point(x):
        x = x.to(torch.float64)
    less = (x.unsqueeze(-1) > x.unsqueeze(-2)).sum(-1)           # [..., r, j]: x_j < x_r
    equal = (x.unsqueeze(-1) == x.unsqueeze(-2)).sum(-1)
    return (1.0 + less + 0.5 * (equal - 1)).to(x.dtype)


def spearmanr_last(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Pearson correlation of the ranks.                   reference: math/spearmanr.py:35-50."""
    return pearsonr_last(rank_last(x), rank_last(y))


def score_fold(y_test: torch.Tensor, y_hat: torch.Tensor, y_train: torch.Tensor, metric: str = 'r2',
               reduce: Sequence[int] = (0,)) -> torch.Tensor:
    """What ``fit_model_`` + ``metrics.score`` compute for one fold: the R2 baseline is the mean of
    the *training* targets over the reduce dims (validator.py:70-89), trials are moved last
    (score.py:95-113)."""
    yt, yh = move_reduce_last(y_test, reduce), move_reduce_last(y_hat, reduce)
    if metric == 'r2':
        yb = move_reduce_last(y_train.mean(tuple(reduce), keepdim=True), reduce)
        return r2_last(yt, yh, yb)
    if metric == 'pearsonr':
        return pearsonr_last(yt, yh)
    raise ValueError(metric)


# --------------------------------------------------------------------------------------------
# TimeDelayed                                           (mvpy/estimators/timedelayed.py:433-535)
# --------------------------------------------------------------------------------------------


def trf_fit(X: torch.Tensor, y: torch.Tensor, window: np.ndarray, alphas: torch.Tensor, fit_intercept=True,
            alpha_per_target=False, patterns=False, solver=ridge_loo_fit) -> Dict[str, torch.Tensor]:
    """Lag-expand, LOO ridge, reshape: coef_ = reshape(f, c, w).flip(-1)   (timedelayed.py:456-466).
    patterns: cov(X_t) coef^T inv(cov(y_t)) reshaped *without* transposing (timedelayed.py:469-494)."""
    f, c, w = y.shape[1], X.shape[1], len(window)
    Xt, yt = delay_design(X, window), flatten_targets(y)
    fit = solver(Xt, yt, alphas, fit_intercept, alpha_per_target)
    out = dict(flat_coef=fit['coef'], intercept=fit['intercept'], alpha=fit['alpha'], losses=fit['losses'],
               coef=fit['coef'].reshape(f, c, w).flip(-1))
    if patterns:
        Sx = torch.cov((Xt - Xt.mean(0, keepdim=True)).T)
        if f > 1:
            Py = torch.linalg.inv(torch.cov((yt - yt.mean(0, keepdim=True)).T))
            pat = Sx @ fit['coef'].T @ Py
        else:
            pat = Sx @ fit['coef'].T * (1.0 / float(yt.shape[0] - 1))
        out['pattern'] = pat.reshape(f, c, w).flip(-1)
    return out


def trf_predict(X: torch.Tensor, window: np.ndarray, flat_coef: torch.Tensor, intercept) -> torch.Tensor:
    """(N, C, T) -> (N, F, T)                            reference: timedelayed.py:522-533."""
    N, _, T = X.shape
    yh = ridge_predict(delay_design(X, window), flat_coef, intercept)
    return yh.reshape(N, T, -1).swapaxes(1, 2)


def cross_val_trf(X: torch.Tensor, y: torch.Tensor, window: np.ndarray, alphas: torch.Tensor, cv: int = 5,
                  scale: bool = True, metrics: Sequence[str] = ('r2',), solver=ridge_loo_fit,
                  folds: Optional[Iterable[int]] = None, scale_dtype: Optional[torch.dtype] = None, scale_device=None,
                  **ridge_kw) -> Dict[str, object]:
    """``cross_val_score(make_pipeline(Scaler().to_torch(), TimeDelayed(...)), X, y, cv=cv)``.

    reference flow: cross_val_score.py:84-99 -> validator.py:321-400 -> fit_model_ (:20-115):
    per fold fit Scaler on X[train] (statistics over trials), fit TimeDelayed, predict X[test] with
    the training statistics, score over trials; scores stacked to (K, F, T).
    ``scale_dtype``: dtype the Scaler step works in (default: X's).  A fp64 check of a fp32 pipeline must scale in fp32
    like the reference does: make_meeg_continuous data holds denormal samples (an AR(1) tail decaying to 1e-45) whose
    variance is exactly 0 in fp32 (scale -> 1, scaler.py:360) but 4e-91 in fp64 (scale 6e-46), a different input.
    ``scale_device``: device the Scaler step runs on (default: X's).  torch's fp32 variance accumulates in double on the CPU
    but in float on CUDA; on the 1e-20-sized samples of the same tails the two differ by 1e-3 relative, so a check against
    the reference's CPU behaviour scales on the CPU."""
    splits = kfold_indices(X.shape[0], cv)
    res = {m: [] for m in metrics}
    alphas_, coefs_, icpts_, losses_ = [], [], [], []
    for k, (tr, te) in enumerate(splits):
        if folds is not None and k not in folds:
            continue
        Xtr, Xte = X[tr], X[te]
        if scale:
            sdt, sdev = scale_dtype or X.dtype, scale_device or X.device
            st = scaler_fit(Xtr.to(device=sdev, dtype=sdt))
            Xtr = scaler_transform(Xtr.to(device=sdev, dtype=sdt), st).to(device=X.device, dtype=X.dtype)
            Xte = scaler_transform(Xte.to(device=sdev, dtype=sdt), st).to(device=X.device, dtype=X.dtype)
        fit = trf_fit(Xtr, y[tr], window, alphas, solver=solver, **ridge_kw)
        yh = trf_predict(Xte, window, fit['flat_coef'], fit['intercept'])
        for m in metrics:
            res[m].append(score_fold(y[te], yh, y[tr], m))
        alphas_.append(fit['alpha'])
        coefs_.append(fit['coef'])
        icpts_.append(fit['intercept'])
        losses_.append(fit['losses'])
    out = {m: torch.stack(v) for m, v in res.items()}
    out.update(alpha=torch.stack(alphas_), coef=torch.stack(coefs_), intercept=icpts_, losses=torch.stack(losses_))
    return out


# --------------------------------------------------------------------------------------------
# hierarchical / Shapley drivers      (mvpy/model_selection/hierarchical.py:17-148, 413-518;
#                                      mvpy/model_selection/shapley.py:64-182)
# --------------------------------------------------------------------------------------------


def group_masks(groups: torch.Tensor) -> Tuple[List[Tuple[int, ...]], torch.Tensor]:
    """All non-empty unions of group rows in ``itertools.combinations`` order, duplicate masks
    dropped keeping the first occurrence.  reference: hierarchical.py:88-131."""
    g = torch.as_tensor(groups).bool()
    ids = range(g.shape[0])
    sets, masks, seen = [], [], set()
    for r in range(1, g.shape[0] + 1):
        for comb in itertools.combinations(ids, r):
            m = g[list(comb)].any(0)
            key = tuple(m.tolist())
            if key in seen:
                continue
            seen.add(key)
            sets.append(comb)
            masks.append(m)
    return sets, torch.stack(masks)


def hierarchical_trf(X, y, groups, window, alphas, cv=5, **kw) -> Dict[str, object]:
    """hierarchical.py:150-207, 413-518: one full cross-validation per feature mask (dim -2)."""
    sets, masks = group_masks(groups)
    scores = [cross_val_trf(X[:, m.to(X.device), :], y, window, alphas, cv=cv, **kw)['r2'] for m in masks]
    return dict(score=torch.stack(scores), sets=sets, masks=masks)


def shapley_trf(X, y, groups, window, alphas, n_permutations=10, cv=5, **kw) -> Dict[str, object]:
    """shapley.py:64-182: per permutation draw ``1 + randperm(G-1)`` then ``randperm(N)`` from the
    global CPU generator (in that order), reshuffle trials, fit the baseline mask (group 0), OR in the
    groups in permuted order, keep score increments, re-sort them to group order."""
    g = torch.as_tensor(groups).bool()
    G = g.shape[0]
    all_scores, orders = [], []
    for _ in range(n_permutations):
        perm = 1 + torch.randperm(G - 1)
        sort_idx = torch.tensor([0] + (perm.argsort() + 1).tolist())
        shuffle = torch.randperm(X.shape[0])
        Xs, ys = X[shuffle], y[shuffle]
        mask = g[0].clone()
        prev = cross_val_trf(Xs[:, mask.to(Xs.device), :], ys, window, alphas, cv=cv, **kw)['r2']
        deltas = [prev]
        for gi in perm.tolist():
            mask = mask | g[gi]
            cur = cross_val_trf(Xs[:, mask.to(Xs.device), :], ys, window, alphas, cv=cv, **kw)['r2']
            deltas.append(cur - prev)
            prev = cur
        all_scores.append(torch.stack(deltas)[sort_idx])
        orders.append(perm)
    return dict(score=torch.stack(all_scores), order=torch.stack(orders))


# --------------------------------------------------------------------------------------------
# one-versus-one probabilities: Wu-Lin pairwise coupling      (mvpy/estimators/classifier.py:527-653)
# --------------------------------------------------------------------------------------------


def pairwise_coupling(p_pair: torch.Tensor, max_iter: int, eps: float, tol: float) -> torch.Tensor:
    """p_pair (N, K, K) -> class probabilities (N, K).  Restates ``_pairwise_coupling_torch`` (classifier.py:573-653) with
    ``_compute_Qr_torch`` (:527-571) written out per index: the reference's in-place adds over repeated index tensors
    (``Q[:, j_i, j_i] += ...``, ``r[:, j_i] += ...`` with j_i = [1, 2, 2, 3, 3, 3, ...]) keep the last write per index, so
    class j's diagonal / right-hand side only see the pairs (j, j - 1) and (K - 1, j); off-diagonal entries are per pair."""
    N, K, _ = p_pair.shape
    dt = p_pair.dtype
    p = torch.full((N, K), 1.0 / K, dtype=dt)
    eye = torch.eye(K, dtype=dt)
    ones = torch.ones((N, K, 1), dtype=dt)
    zeros = torch.zeros((N, 1, 1), dtype=dt)

    def terms(j, k):
        pij, pji = p_pair[:, j, k], p_pair[:, k, j]
        w = 1.0 / torch.clamp(pij * pji, min=eps)
        denom = p[:, j] + p[:, k] + eps
        return pij, pji, w, denom

    for _ in range(max_iter):
        Q = torch.zeros((N, K, K), dtype=dt)
        r = torch.zeros((N, K), dtype=dt)
        for j in range(1, K):
            pij, pji, w, denom = terms(j, j - 1)
            Q[:, j, j] = w * pji / denom ** 2
            r[:, j] = w * (pij - p[:, j] / denom)
        for k in range(K - 1):
            pij, pji, w, denom = terms(K - 1, k)
            Q[:, k, k] += w * pij / denom ** 2
            r[:, k] += w * (pji - p[:, k] / denom)
        for j in range(1, K):
            for k in range(j):
                pij, pji, w, denom = terms(j, k)
                Q[:, j, k] -= w * pji / denom ** 2
                Q[:, k, j] -= w * pij / denom ** 2
        A = Q + 1e-9 * eye
        kkt = torch.cat([torch.cat([A, ones], 2), torch.cat([ones.transpose(1, 2), zeros], 2)], 1)
        delta = torch.linalg.solve(kkt, torch.cat([r, zeros[:, :, 0]], 1).unsqueeze(2)).squeeze(2)[:, :K]
        p_n = torch.clamp(p + delta, min=0)
        s = p_n.sum(1, keepdim=True)
        p_n = torch.where(s > eps, p_n / s, torch.full_like(p_n, 1.0 / K))
        done = torch.max(torch.sum(torch.abs(p_n - p), dim=1)) < tol
        p = p_n
        if done:
            break
    return p


# --------------------------------------------------------------------------------------------
# decoders / encoders / sliding      (mvpy/estimators/ridgedecoder.py:244-310,
#                                     mvpy/estimators/ridgeencoder.py:273-382, sliding.py:575-622)
# --------------------------------------------------------------------------------------------


def decoder_fit(X: torch.Tensor, y: torch.Tensor, alphas, fit_intercept=True, alpha_per_target=True,
                solver=ridge_loo_fit) -> Dict[str, torch.Tensor]:
    """RidgeDecoder: LOO ridge (per-target alpha by default) + Haufe pattern
    cov(X) coef^T pinv(cov(y)).  reference: ridgedecoder.py:244-285.  (The reference's public
    ``RidgeDecoder`` constructor is broken at this commit -- SURVEY.md section 8c defect 1 -- so this
    restates the intended computation: ``_RidgeCV_torch`` + the pattern formula.)"""
    if y.ndim == 1:
        y = y[:, None]
    fit = solver(X, y, alphas, fit_intercept, alpha_per_target)
    Sx = torch.cov((X - X.mean(0, keepdim=True)).T)
    if y.shape[1] > 1:
        Py = torch.linalg.pinv(torch.cov((y - y.mean(0, keepdim=True)).T))
        pat = Sx @ fit['coef'].T @ Py
    else:
        pat = Sx @ fit['coef'].T * (1.0 / float(y.shape[0] - 1))
    return dict(coef=fit['coef'], intercept=fit['intercept'], alpha=fit['alpha'], pattern=pat, losses=fit['losses'])


def sliding_decoder_fit(X: torch.Tensor, y: torch.Tensor, alphas, **kw) -> List[Dict[str, torch.Tensor]]:
    """Sliding(RidgeDecoder, dims=-1): one independent fit per slice of the last dim; a singleton last
    dim on either side is broadcast.  reference: sliding.py:596-618."""
    Tx, Ty = X.shape[-1], y.shape[-1]
    n = max(Tx, Ty)
    return [decoder_fit(X[..., i if Tx > 1 else 0], y[..., i if Ty > 1 else 0], alphas, **kw) for i in range(n)]


def encoder_expand(X: torch.Tensor) -> torch.Tensor:
    """(N, F, T) -> block design (N*T, T*F): row i*T+j holds X[i, :, j] in columns j*F..(j+1)*F-1.
    reference: ridgeencoder.py:292-304 (there as a Python double loop)."""
    N, F, T = X.shape
    out = torch.zeros(N, T, T, F, dtype=X.dtype)
    idx = torch.arange(T)
    out[:, idx, idx, :] = X.permute(0, 2, 1)
    return out.reshape(N * T, T * F)


def encoder_fit(X: torch.Tensor, y: torch.Tensor, alphas, solver=ridge_loo_fit, **kw) -> Dict[str, torch.Tensor]:
    """RidgeEncoder.  2-D: plain LOO ridge.  3-D: expanded design, one shared alpha/intercept,
    coef reshaped (d, t, f) -> swap -> (d, f, t).  reference: ridgeencoder.py:307-345."""
    if X.ndim == 3 and y.ndim == 3:
        N, F, T = X.shape
        d = y.shape[1]
        fit = solver(encoder_expand(X), flatten_targets(y), alphas, **kw)
        return dict(coef=fit['coef'].reshape(d, T, F).swapaxes(1, 2), flat_coef=fit['coef'],
                    intercept=fit['intercept'], alpha=fit['alpha'], expanded=True)
    fit = solver(X, y, alphas, **kw)
    return dict(coef=fit['coef'], flat_coef=fit['coef'], intercept=fit['intercept'], alpha=fit['alpha'], expanded=False)


def encoder_predict(X: torch.Tensor, fit: Dict[str, torch.Tensor]) -> torch.Tensor:
    """reference: ridgeencoder.py:347-382."""
    if fit['expanded']:
        N, _, T = X.shape
        yh = ridge_predict(encoder_expand(X), fit['flat_coef'], fit['intercept'])
        return yh.reshape(N, T, -1).swapaxes(1, 2)
    return ridge_predict(X, fit['flat_coef'], fit['intercept'])


# --------------------------------------------------------------------------------------------
# ReceptiveField                  (mvpy/estimators/receptivefield.py:779-1556)
# --------------------------------------------------------------------------------------------
# Lag k of the filter stands for sample lag s_k = s_min + k, k = 0..W-1, W = s_max - s_min; the model is
#   y[n, f, t] = sum_c sum_k coef[f, c, k] * X[n, c, t - s_k]   (X is zero outside 0..T-1)   (:1434-1488).
# The reference builds the normal equations per epoch from FFT correlations; the same numbers are written
# here as direct lag sums in the dtype of the inputs (use fp64).


def rf_lags(t_min: float, t_max: float, fs: float) -> Tuple[int, int]:
    """(s_min, s_max) with s_max exclusive (receptivefield.py:955-957)."""
    return int(t_min * fs), int(t_max * fs) + 1


def _lagged(x: torch.Tensor, lag: int) -> torch.Tensor:
    """x (..., T) -> z with z[t] = x[t - lag], zero outside."""
    T = x.shape[-1]
    z = torch.zeros_like(x)
    if lag >= 0:
        if lag < T:
            z[..., lag:] = x[..., :T - lag]
    elif -lag < T:
        z[..., :T + lag] = x[..., -lag:]
    return z


def _full_xcorr(a: torch.Tensor, b: torch.Tensor, tau: int) -> torch.Tensor:
    """sum_t a[t + tau] b[t] over every t where both samples exist (the zero-padded FFT correlation, :1040-1055)."""
    T = a.shape[-1]
    if tau >= 0:
        return (a[tau:] * b[:T - tau]).sum() if tau < T else a.new_zeros(())
    return (a[:T + tau] * b[-tau:]).sum() if -tau < T else a.new_zeros(())


def _diag_cumulative(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """out[m, n] = sum_{d=0..min(m,n)} a[m-d] b[n-d]: the outer product accumulated down its diagonals (:810-830)."""
    L = a.shape[0]
    out = a.new_zeros((L, L))
    for m in range(L):
        for n in range(L):
            d = torch.arange(min(m, n) + 1)
            out[m, n] = (a[m - d] * b[n - d]).sum()
    return out


def rf_corrs(X: torch.Tensor, y: torch.Tensor, s_min: int, s_max: int, fit_intercept: bool = True,
             edge_correction: bool = True):
    """Per-epoch auto- and cross-correlation blocks (receptivefield.py:976-1118).

    Returns cov (N, C*W, C*W), xty (N, C*W, F) [row c*W+k], X_offset (C,) / y_offset (F,) (zeros without intercept).
    Reproduced as the reference computes them, including two things a textbook version would do differently:
      * the block below the diagonal, (j, i) with j > i, receives the SAME W x W lag block as (i, j), not its
        transpose (:1097-1103), so cov is not symmetric for C > 1;
      * the correction for rows before the epoch start (negative lags) accumulates the first |s_min| samples in
        reversed order without flipping the result back (:801-807 with the flipped epoch passed at :1083).
    """
    N, C, T = X.shape
    F = y.shape[1]
    W = s_max - s_min
    if fit_intercept:
        x_off, y_off = X.mean((0, 2)), y.mean((0, 2))
        X, y = X - x_off[None, :, None], y - y_off[None, :, None]
    else:
        x_off, y_off = X.new_zeros(C), y.new_zeros(F)
    cov = X.new_zeros((N, C * W, C * W))
    xty = X.new_zeros((N, C * W, F))
    for n in range(N):
        flipped = X[n].flip(-1)
        for i in range(C):
            for j in range(i, C):
                blk = X.new_zeros((W, W))
                for a in range(W):
                    for b in range(W):
                        blk[a, b] = _full_xcorr(X[n, i], X[n, j], b - a)
                if edge_correction:
                    if s_max > 0:                       # rows past the epoch end (positive lags)
                        tail = _diag_cumulative(flipped[i, :s_max - 1], flipped[j, :s_max - 1])
                        if s_min > 0:
                            tail = tail[s_min - 1:, s_min - 1:]
                        o = max(-s_min + 1, 0)
                        blk[o:, o:] -= tail
                    if s_min < 0:                       # rows before the epoch start (negative lags)
                        head = _diag_cumulative(flipped[i, s_min:], flipped[j, s_min:])
                        if s_max < 0:
                            head = head[:s_max, :s_max]
                        blk[:-s_min, :-s_min] -= head
                cov[n, i * W:(i + 1) * W, j * W:(j + 1) * W] += blk
                if i != j:
                    cov[n, j * W:(j + 1) * W, i * W:(i + 1) * W] += blk
            for k in range(W):
                xty[n, i * W + k] = (y[n] * _lagged(X[n, i], s_min + k)[None, :]).sum(-1)
    return cov, xty, x_off, y_off


def rf_regulariser(C: int, W: int, reg_type='ridge') -> torch.Tensor:
    """Penalty matrix (receptivefield.py:1120-1183): identity for ridge/ridge; otherwise the sum of a second-
    difference (free-end) operator along lags within each feature and / or along features within each lag --
    without the identity when only the feature direction is 'laplacian'."""
    if isinstance(reg_type, str):
        reg_type = (reg_type,) * 2
    if len(reg_type) != 2:
        raise ValueError(f'reg_type must be of length two, but got {len(reg_type)}.')
    for r in reg_type:
        if r not in ('ridge', 'laplacian'):
            raise ValueError(f'Unknown reg_type. Expected ridge or laplacian, but got {r}.')

    def second_difference(m: int) -> torch.Tensor:
        L = torch.zeros((m, m))
        for r in range(m - 1):
            L[r, r] += 1; L[r + 1, r + 1] += 1; L[r, r + 1] -= 1; L[r + 1, r] -= 1
        return L

    over_time = reg_type[0] == 'laplacian' and W > 1
    over_feat = reg_type[1] == 'laplacian' and C > 1
    if not over_time and not over_feat:
        return torch.eye(C * W)
    reg = torch.zeros((C * W, C * W))
    if over_time:
        reg += torch.kron(torch.eye(C), second_difference(W))
    if over_feat:
        reg += torch.kron(second_difference(C), torch.eye(W))
    return reg


def rf_solve(cov: torch.Tensor, xty: torch.Tensor, alpha, reg: torch.Tensor, C: int, W: int) -> torch.Tensor:
    """Sum the epochs, add alpha * reg, solve; (F, C, W) coefficients (receptivefield.py:1185-1244)."""
    G, b = cov.sum(0), xty.sum(0)
    w = torch.linalg.solve(G + float(alpha) * reg.to(G.dtype), b)
    return w.T.reshape(b.shape[1], C, W)


def rf_predict(X: torch.Tensor, coef: torch.Tensor, intercept, s_min: int) -> torch.Tensor:
    """(N, F, T) predictions of the lagged model above, plus intercept (receptivefield.py:1434-1488)."""
    Fn, C, W = coef.shape
    out = X.new_zeros((X.shape[0], Fn, X.shape[-1]))
    for k in range(W):
        out += torch.einsum('fc,nct->nft', coef[:, :, k], _lagged(X, s_min + k))
    if isinstance(intercept, torch.Tensor):
        out = out + intercept.reshape(1, -1, 1)
    return out


def rf_fit(X: torch.Tensor, y: torch.Tensor, t_min: float, t_max: float, fs: float, alpha, reg_type='ridge', reg_cv=5,
           fit_intercept: bool = True, edge_correction: bool = True, patterns: bool = False) -> Dict[str, object]:
    """ReceptiveField.fit (receptivefield.py:1322-1432): one solve for a float alpha; a tensor of alphas is scored
    by k-fold cross-validation over epochs (mean Pearson correlation across the held-out epochs, averaged over
    time, channels and folds; first maximum wins) or, for reg_cv='LOO', handed to the LOO ridge solver with the
    summed correlation matrix as its "data" (:1213-1221)."""
    s_min, s_max = rf_lags(t_min, t_max, fs)
    W = s_max - s_min
    N, C, T = X.shape
    F = y.shape[1]
    cov, xty, x_off, y_off = rf_corrs(X, y, s_min, s_max, fit_intercept, edge_correction)
    out: Dict[str, object] = dict(cov=cov, s_min=s_min, s_max=s_max)

    def intercept_of(coef):
        return (y_off - x_off @ coef.sum(-1).T) if fit_intercept else 0.0

    if reg_cv == 'LOO':
        fit = ridge_loo_fit(cov.sum(0), xty.sum(0), alpha.to(X.dtype), fit_intercept=True, alpha_per_target=False)
        coef, best_alpha = fit['coef'].reshape(F, C, W), fit['alpha']
    else:
        reg = rf_regulariser(C, W, reg_type)
        if isinstance(alpha, float):
            best_alpha = alpha
        else:
            folds = kfold_indices(N, reg_cv) if isinstance(reg_cv, int) else list(reg_cv)
            means = []
            for a in alpha:
                per_fold = []
                for train, test in folds:
                    c = rf_solve(cov[train], xty[train], a, reg, C, W)
                    y_h = rf_predict(X[test], c, intercept_of(c), s_min).float()
                    per_fold.append(pearsonr_last(y_h.permute(2, 1, 0), y[test].permute(2, 1, 0)).mean().item())
                means.append(torch.mean(torch.tensor(per_fold)).item())
            out['cv_scores'] = torch.tensor(means)
            best_alpha = alpha[torch.argmax(out['cv_scores'])]
        coef = rf_solve(cov, xty, best_alpha, reg, C, W)
    out.update(coef=coef, alpha=best_alpha, intercept=intercept_of(coef))
    if patterns:
        S_X = cov.sum(0) / float(T * N - 1)
        if F > 1:
            z = y.swapaxes(1, 2).reshape(-1, F)
            z = z - z.mean(0, keepdim=True)
            out['pattern'] = (S_X @ (coef.reshape(C * W, F) @ torch.linalg.pinv(torch.cov(z.T)))).reshape(coef.shape)
        else:
            out['pattern'] = (S_X @ (coef.reshape(C * W, F) / float(T * N - 1))).reshape(coef.shape)
    return out
