"""Hierarchical (all-subsets) predictor-set scoring -- drop-in for ``mvpy.model_selection.Hierarchical``
(mvpy/model_selection/hierarchical.py:17-518).

Host side: enumerate all non-empty unions of group rows in ``itertools.combinations`` order, drop
duplicate masks keeping the first occurrence, run one full cross-validation per mask.  The (set)
units are independent: under ``torch.distributed`` they are sharded over the GPUs and the scores
all-gathered once (``_dist``)."""
from __future__ import annotations

import warnings
from itertools import chain, combinations
from typing import Dict, List, Optional, Tuple, Union

import torch

from .. import _dist, _lib
from ..crossvalidation import Validator
from ..crossvalidation.validator import _move_tensors


def prepare_groups_(X: torch.Tensor, groups, dim: Optional[int], on_underspecified: str = 'raise'):
    """Shared argument handling (hierarchical.py:52-86, shapley.py:378-418): default feature dim (-2 for
    3-D X, else -1), identity groups by default, list -> tensor, coverage check, bool mask."""
    if dim is None:
        dim = -2 if X.ndim > 2 else -1
    if groups is None:
        groups = torch.eye(X.shape[dim], dtype=torch.bool, device=X.device)
    else:
        if isinstance(groups, list):
            groups = torch.tensor(groups, dtype=torch.long, device=X.device)
        if groups.shape[1] != X.shape[dim]:
            raise ValueError(f'Expected `groups` to be of shape (n_groups, n_features) matching n_features in `X`, '
                             f'but got groups={groups.shape} and X={X.shape} with dim={dim}.')
        if not (groups.sum(0) > 0).all():
            if on_underspecified == 'warn':
                warnings.warn(f'Supplied `groups` do not include all available predictors. Counted {groups.sum(0)} instances.')
            elif on_underspecified != 'ignore':
                raise ValueError('When supplying `groups`, each predictor must appear in at least one grouping.')
        groups = groups.to(torch.bool)
    return dim, groups


def check_dims_and_groups_(X: torch.Tensor, y: torch.Tensor, groups=None, dim: Optional[int] = None,
                           on_underspecified: str = 'raise'):
    """-> (dim, group_ids, sets, masks); sets in combinations order by subset size, duplicate masks removed
    keeping the first occurrence (hierarchical.py:88-131)."""
    dim, groups = prepare_groups_(X, groups, dim, on_underspecified)
    n_groups = groups.shape[0]
    group_ids = list(range(n_groups))
    sets, masks, seen = [], [], set()
    g_host = groups.cpu()
    for comb in chain.from_iterable(combinations(group_ids, r) for r in range(1, n_groups + 1)):
        m = g_host[list(comb)].any(0)
        key = tuple(m.tolist())
        if key in seen:
            continue
        seen.add(key)
        sets.append(comb)
        masks.append(m)
    masks = torch.stack(masks, dim=0).to(X.device)
    return dim, group_ids, sets, masks, groups


def fit_validator_(validator: Validator, X: torch.Tensor, y: torch.Tensor, mask: torch.Tensor, dim: int,
                   shared: Optional[dict] = None, split_like=None) -> Validator:
    """One cross-validation on the predictors selected by ``mask`` (hierarchical.py:150-207).  ``shared``: a dict owned
    by the caller and handed to every predictor set of the same (X, y): estimators that can (``TimeDelayed``) keep each
    fold's full-predictor Gram / cross-moment in it and slice them per set instead of recomputing (``_lib.fold_ctx``)."""
    idx = torch.nonzero(mask.to(X.device), as_tuple=False).squeeze(-1)
    X_i = torch.index_select(X, dim=dim, index=idx)
    v = validator.clone()
    _lib.fold_ctx.parent = None if shared is None else dict(X=X, dim=dim, idx=idx, shared=shared)
    try:
        v.fit(X_i, y, **({} if split_like is None else {'_split_like': split_like}))
    finally:
        _lib.fold_ctx.parent = None
    return v


def _stack_scores(scores: List, multi_names):
    if multi_names is not None:
        return {name: torch.stack([s[name] for s in scores], dim=0) for name in multi_names}
    return torch.stack(scores, dim=0)


class Hierarchical:
    """Attributes: ``validator_`` (fitted per set; ``None`` if fitted on another rank), ``score_``
    ((n_sets, n_folds, ...)), ``mask_, group_, group_id_, set_``."""

    def __init__(self, validator: Validator, n_jobs: Optional[int] = None, verbose: Union[int, bool] = False,
                 on_underspecified: str = 'raise'):
        self.validator = validator
        self.n_jobs = n_jobs
        self.verbose = verbose
        self.on_underspecified = on_underspecified
        self.validator_ = []
        self.score_ = []
        self.mask_ = []
        self.group_ = []
        self.group_id_ = []
        self.set_ = []

    def fit(self, X: torch.Tensor, y: torch.Tensor, groups=None, dim: Optional[int] = None) -> 'Hierarchical':
        if not (isinstance(X, torch.Tensor) and isinstance(y, torch.Tensor)):
            raise ValueError(f'Both `X` and `y` must be of type torch.Tensor, but got {type(X)} and {type(y)}.')
        dim, group_ids, sets, masks, groups = check_dims_and_groups_(X, y, groups=groups, dim=dim,
                                                                     on_underspecified=self.on_underspecified)
        out_dev = X.device
        multi = [m.name for m in self.validator.metric] if (self.validator.metric is not None and len(self.validator.metric) > 1) else None
        shard = _dist.sharding_active()
        Xd, yd = _lib.stage(X), _lib.stage(y)          # one host->device copy for all sets
        # folds are drawn on the device of the user's tensors, not on the staging device (fold membership bit-exact with the
        # reference for shuffling splitters)
        split_like = (X, y) if Xd.device != X.device else None
        if shard:
            self._fit_sharded(Xd, yd, masks, dim, multi, out_dev, split_like)
        else:
            self.validator_, local = [], {}
            # per-fold statistics of the full predictor set, shared by all sets (pointless for a single set)
            shared: Optional[dict] = {} if len(masks) > 1 else None
            with _dist.sharded_region():
                for i, mask in enumerate(masks):
                    v = fit_validator_(self.validator, Xd, yd, mask, dim, shared, split_like)
                    local[i] = v.score_
                    self.validator_.append(_move_tensors(v, out_dev) if Xd.device != out_dev else v)
            if multi is not None:
                self.score_ = {name: torch.stack([local[i][name] for i in range(len(masks))], dim=0).to(out_dev) for name in multi}
            else:
                self.score_ = torch.stack([local[i] for i in range(len(masks))], dim=0).to(out_dev)
        self.mask_ = [masks[i] for i in range(len(masks))]
        self.set_ = list(sets)
        self.group_ = groups
        self.group_id_ = group_ids
        return self

    def _fit_sharded(self, Xd: torch.Tensor, yd: torch.Tensor, masks: torch.Tensor, dim: int, multi, out_dev, split_like=None) -> None:
        """Multi-GPU: the (set, fold) units -- not whole sets, whose costs differ by the cube of their column counts -- are
        spread over the ranks longest-first (``_dist.lpt_assign``, the joblib site hierarchical.py:468 hands out whole sets).
        Every rank walks its units fold by fold so that the fold's full-predictor statistics are shared by its sets; the
        score shards are combined once.  ``validator_[i].model_[k]`` is None for folds fitted on another rank."""
        from ..crossvalidation.validator import fit_model_
        n_sets = len(masks)
        base = self.validator
        folds = [(tr.cpu(), te.cpu()) for tr, te in base.cv_.split(*(split_like if split_like is not None else (Xd, yd)))]
        folds = _dist.broadcast_object(folds, src=0)                  # shuffling splitters: rank 0's draw is authoritative
        K = len(folds)
        n_feat = [int(m.sum()) for m in masks]
        costs = [float(n_feat[i]) ** 2.5 for i in range(n_sets) for _ in range(K)]          # between the p^2 n and p^3 terms
        owner = _dist.lpt_assign(costs, _dist.world()[1])
        rank = _dist.world()[0]
        mine = [u for u in range(n_sets * K) if owner[u] == rank]
        mine.sort(key=lambda u: (u % K, u // K))                      # fold-major: one set of shared statistics alive at a time
        vals = [base.clone() for _ in range(n_sets)]
        for v in vals:
            v.model_, v.metric_, v.test_ = [None] * K, [None] * K, [te for _, te in folds]
        from ..crossvalidation.validator import fold_key, _lanes_ok
        from .. import _streams
        local, shared = {}, {}                                        # `shared` is bounded by the estimator (oldest fold evicted)
        idxs = [torch.nonzero(m.to(Xd.device), as_tuple=False).squeeze(-1) for m in masks]
        keys = [fold_key(tr) for tr, _ in folds]

        def one(u: int):
            i, k = u // K, u % K
            X_i = torch.index_select(Xd, dim=dim, index=idxs[i])
            _lib.fold_ctx.parent = dict(X=Xd, dim=dim, idx=idxs[i], shared=shared)
            try:
                return fit_model_(base.model, folds[k][0].to(Xd.device), folds[k][1].to(Xd.device), X_i, y=yd, metric=base.metric,
                                  train_key=keys[k])
            finally:
                _lib.fold_ctx.parent = None
        with _dist.sharded_region():
            # units of one fold share its statistics and stay on one stream lane, in order; different folds overlap
            if Xd.is_cuda and _lanes_ok(base.model):
                done = _streams.run_on_lanes([(lambda u=u: one(u)) for u in mine], Xd.device, lane_keys=[u % K for u in mine])
            else:
                done = [one(u) for u in mine]
        for u, r in zip(mine, done):
            i, k = u // K, u % K
            local[u] = r['score_']
            vals[i].model_[k] = _move_tensors(r['model'], out_dev)
            vals[i].metric_[k] = _move_tensors(r['metric'], out_dev)
        names = multi if multi is not None else [None]
        scores = {}
        for name in names:
            per = _dist.gather_assigned({u: (s[name] if name is not None else s) for u, s in local.items()}, n_sets * K)
            scores[name] = torch.stack([torch.stack(per[i * K:(i + 1) * K], dim=0) for i in range(n_sets)], dim=0).to(out_dev)
        for i, v in enumerate(vals):
            v.score_ = {name: scores[name][i] for name in multi} if multi is not None else scores[None][i]
        self.validator_ = vals
        self.score_ = {name: scores[name] for name in multi} if multi is not None else scores[None]
