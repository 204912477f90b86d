"""Batched execution of many small, identically shaped ridge fits (``Sliding`` over ``RidgeDecoder`` / ``RidgeCV``).

The reference fits one estimator clone per slice in a Python loop (mvpy/estimators/sliding.py:575-622).  Here the slices are
the batch index of ``mvb_dense_ridge_batched`` (csrc/dense_batched.cuh): all statistics, one Cholesky factor per
(slice, alpha), the leverages and fitted values on the tensor cores, losses, alpha selection and coefficients in ~20 launches
for the whole stack, reading the user's (trials, channels, time) array through its strides.  The result is unpacked into the
same per-slice estimator objects the loop would have produced (``estimators_``), so everything downstream is unchanged.

Slices whose factorisation met a small pivot (``status != 0``: a rank-deficient or badly scaled slice at the smallest
alphas) are refitted one by one on the general route, which has the conditioning safeguards (block Lanczos + guarded
sweeps / fp64 for wide problems).  ``MVB_SLIDING_BATCHED=0`` switches this executor off (CUDA-graph replay or the eager loop then run).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib


def _unwrap(proto):
    """(kind, alphas, fit_intercept, alpha_per_target) of a ridge prototype this executor understands, else None."""
    from .estimators.ridgecv import _RidgeCV_torch
    from .estimators.ridgedecoder import _RidgeDecoder_torch
    if type(proto) is _RidgeDecoder_torch:
        return 'decoder', proto.alpha, bool(proto.fit_intercept), bool(proto.alpha_per_target)
    if type(proto) is _RidgeCV_torch:
        return 'ridgecv', proto.alphas, bool(proto.fit_intercept), bool(proto.alpha_per_target)
    return None


def _patterns(gram: torch.Tensor, coef: torch.Tensor, Y3: torch.Tensor, n: int) -> torch.Tensor:
    """Haufe patterns of every slice, cov(X) coef^T pinv(cov(y)) (ridgedecoder.py:267-285): (B, p, F).  cov(X) is the centred
    Gram the fit built; the (F x F) target covariances are inverted through the library's batched Jacobi eigensolver."""
    B, F, p = coef.shape
    Sx = gram / float(n - 1)
    T1 = torch.bmm(Sx, coef.transpose(1, 2))                                   # (B, p, F)
    if F == 1:
        return T1 * (1.0 / float(n - 1))
    yc = Y3 - Y3.mean(0, keepdim=True)                                         # (n, F, B)
    Sy = torch.einsum('nfb,ngb->bfg', yc, yc) / float(n - 1)
    Vt, lam = _lib.eigh(Sy.contiguous().clone())
    keep = lam > lam.max(dim=1, keepdim=True).values * (F * torch.finfo(Sy.dtype).eps)
    inv = torch.where(keep, 1.0 / torch.where(keep, lam, torch.ones_like(lam)), torch.zeros_like(lam))
    Py = torch.einsum('bkf,bk,bkg->bfg', Vt, inv, Vt)
    return torch.bmm(T1, Py)


def fit_slices(proto, Xm: torch.Tensor, ym: torch.Tensor, pairs: Sequence[Tuple[int, int]]) -> Optional[List]:
    """Fit ``proto.clone()`` on (Xm[..., i], ym[..., j]) for every pair; returns the fitted estimators, or None when this
    executor does not take the problem (the caller falls back to graph replay / the eager loop)."""
    if os.environ.get('MVB_SLIDING_BATCHED', '1') == '0':
        return None
    info = _unwrap(proto)
    if info is None or not (isinstance(Xm, torch.Tensor) and isinstance(ym, torch.Tensor)):
        return None
    kind, alphas, fit_intercept, per_target = info
    if not isinstance(alphas, torch.Tensor) or Xm.ndim != 3 or ym.ndim != 3 or Xm.shape[0] != ym.shape[0]:
        return None
    if Xm.dtype != torch.float32 or ym.dtype != torch.float32 or len(pairs) < 2:
        return None
    if not torch.cuda.is_available() or torch.cuda.is_current_stream_capturing():
        return None
    n, p, F, A = Xm.shape[0], Xm.shape[1], ym.shape[1], int(alphas.numel())
    if not _lib.dense_ridge_batched_ok(n, p, F, A, Xm.dtype):
        return None
    B = len(pairs)
    xi, yi = [i for i, _ in pairs], [j for _, j in pairs]
    out_dev = Xm.device
    Xd, yd = _lib.to_device(Xm), _lib.to_device(ym)           # host inputs: one copy of the whole array, not one per slice

    def stack_view(t: torch.Tensor, idx: List[int]) -> Optional[torch.Tensor]:
        if idx == list(range(t.shape[-1])):
            return t
        if len(set(idx)) == 1:                                  # a singleton sliding dim is broadcast (sliding.py:606-611)
            return t[..., idx[0]:idx[0] + 1].expand(-1, -1, B)
        return None
    X3, Y3 = stack_view(Xd, xi), stack_view(yd, yi)
    if X3 is None or Y3 is None:
        return None
    al = alphas.to(Xm.dtype)
    res = _lib.dense_ridge_batched(X3, Y3, None, _lib.to_device(al), fit_intercept, per_target, want_loss=True,
                                   want_gram=(kind == 'decoder'))
    coef, icpt, alpha, loss = res['coef'], res['intercept'], res['alpha'], res['loss']
    pattern = _patterns(res['gram'], coef, Y3, n) if kind == 'decoder' else None
    bad = torch.nonzero(res['status']).reshape(-1).tolist()      # one read-back per stack of slices
    coef_o, icpt_o, alpha_o, loss_o = coef.to(out_dev), icpt.to(out_dev), alpha.to(out_dev), loss.to(out_dev)
    pattern_o = None if pattern is None else pattern.to(out_dev)
    al_o = al.to(out_dev)
    stack = dict(kind=kind, per_target=per_target, fit_intercept=fit_intercept, alphas=al_o, coef=coef_o, alpha=alpha_o, loss=loss_o,
                 icpt=icpt_o, pattern=pattern_o)
    fitted = [proto.clone() for _ in range(B)]
    assign_views(fitted, stack)
    for b in bad:                                                # rank-deficient / badly scaled slices: the guarded general route
        fitted[b] = proto.clone().fit(Xm[..., xi[b]], ym[..., yi[b]])
    # the stacked parameters stay with the caller so that predict needs no re-stacking (only when every slice came from here);
    # `stack` (all per-slice attributes as stacked arrays) lets a driver move the whole list to another device with a few copies
    stacked = None if bad else (coef, icpt if fit_intercept else None)
    return fitted, stacked, (None if bad else stack)


def assign_views(estimators: Sequence, stack: dict) -> None:
    """Point the fitted attributes of the per-slice estimators at rows of the stacked arrays (a handful of unbind calls instead of
    thousands of Python-level indexing operations)."""
    kind, per_target, fit_intercept = stack['kind'], stack['per_target'], stack['fit_intercept']
    coef_v, alpha_v, loss_v = stack['coef'].unbind(0), stack['alpha'].unbind(0), stack['loss'].unbind(0)
    icpt_v = stack['icpt'][:, None, :].unbind(0) if fit_intercept else None
    pattern_v = stack['pattern'].unbind(0) if stack['pattern'] is not None else None
    for b, est in enumerate(estimators):
        inner = est.estimator if kind == 'decoder' else est
        inner.alphas = stack['alphas']                           # ridgecv.py:125
        inner.coef_ = coef_v[b]
        inner.alpha_ = alpha_v[b]                                # (F,) per target, 0-d when shared
        inner.loss_ = loss_v[b]
        inner.intercept_ = icpt_v[b] if fit_intercept else 0.0
        inner._gram = None
        if kind == 'decoder':
            est.coef_, est.intercept_, est.alpha_ = inner.coef_, inner.intercept_, inner.alpha_
            est.pattern_ = pattern_v[b]


def move_stack(estimators: Sequence, stack: dict, device: torch.device) -> dict:
    """The per-slice estimators of a batched fit on another device: the stacked arrays move (one copy each), the views follow."""
    new = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in stack.items()}
    assign_views(estimators, new)
    return new


def stacked_parameters(estimators: Sequence) -> Optional[Tuple[torch.Tensor, Optional[torch.Tensor]]]:
    """(coef (B, F, p), intercept (B, F) | None) of a list of fitted ridge estimators, or None if they are anything else."""
    if not estimators or _unwrap(estimators[0]) is None:
        return None
    kind = _unwrap(estimators[0])[0]
    inner = [e.estimator if kind == 'decoder' else e for e in estimators]
    if any(e.coef_ is None or e.coef_.ndim != 2 for e in inner):
        return None
    coef = torch.stack([e.coef_ for e in inner])
    has_icpt = isinstance(inner[0].intercept_, torch.Tensor)
    icpt = torch.stack([e.intercept_.reshape(-1) for e in inner]) if has_icpt else None
    return coef, icpt


def predict_slices(estimators: Sequence, Xm: torch.Tensor, stacked=None) -> Optional[torch.Tensor]:
    """Predictions of one fitted ridge estimator per slice of Xm (N, p, B) -> (N, F, B) in one launch
    (``mvb_dense_predict_batched``; the reference loops, sliding.py:770-831), or None when not applicable."""
    if os.environ.get('MVB_SLIDING_BATCHED', '1') == '0' or not isinstance(Xm, torch.Tensor):
        return None
    if Xm.ndim != 3 or Xm.dtype != torch.float32 or len(estimators) != Xm.shape[-1] or not torch.cuda.is_available():
        return None
    params = stacked if stacked is not None else stacked_parameters(estimators)
    if params is None:
        return None
    coef, icpt = params
    if coef.dtype != torch.float32 or coef.shape[2] != Xm.shape[1]:
        return None
    out_dev = Xm.device
    yh = _lib.dense_predict_batched(_lib.to_device(Xm), _lib.to_device(coef).contiguous(), None if icpt is None else _lib.to_device(icpt).contiguous())
    return yh.to(out_dev)
