"""Permutation Shapley values over predictor groups -- drop-in for ``mvpy.model_selection.Shapley``
(mvpy/model_selection/shapley.py:17-490).

Per permutation the reference draws ``1 + randperm(G-1)`` (group order) and ``randperm(N)`` (trial
order) from the global RNG of ``X.device`` -- in that order -- reshuffles the trials, cross-validates
the baseline mask (group 0) and then the nested masks obtained by OR-ing in the groups in permuted
order, stores score increments and re-sorts them to group order.  The RNG calls are kept as torch
calls in the reference's order so permutation orders and fold memberships are bit-exact (SURVEY.md
H6).  Permutations are independent: under ``torch.distributed`` rank 0 draws all of them first and
broadcasts, every rank processes its share, scores are all-gathered once."""
from __future__ import annotations

from typing import Dict, List, Optional, Union

import torch

from .. import _dist, _lib
from ..crossvalidation import Validator
from ..crossvalidation.validator import _move_tensors
from .hierarchical import fit_validator_, prepare_groups_


def draw_permutation_(n_groups: int, n_trials: int, device) -> Dict[str, torch.Tensor]:
    """The two RNG draws of shapley.py:77,82 (global generator of ``device``), plus the re-sort index."""
    perm = 1 + torch.randperm(n_groups - 1, device=device)
    sort = torch.tensor([0] + (perm.argsort() + 1).cpu().tolist(), dtype=torch.long, device=device)
    data = torch.randperm(n_trials, device=device)
    return dict(perm=perm, sort=sort, data=data)


def fit_permutation_(validator: Validator, X: torch.Tensor, y: torch.Tensor, groups: torch.Tensor, dim: int,
                     draw: Dict[str, torch.Tensor], keep_validators: bool = True, orig=None) -> Dict:
    """shapley.py:64-182 for one permutation, given its RNG draws.  X, y live on the compute device; ``orig`` = the user's
    (X, y) when they live elsewhere: the folds of the reshuffled trials are then drawn on that device."""
    perm, sort, data = draw['perm'], draw['sort'], draw['data'].to(X.device)
    Xs, ys = X[data], y[data]
    split_like = None
    if orig is not None:
        d0 = draw['data'].to(orig[0].device)
        split_like = (orig[0][d0], orig[1][d0])
    mask = groups[0, :].clone()
    shared: dict = {}                 # per-fold full-predictor statistics, shared by the nested sets of this permutation
    v = fit_validator_(validator, Xs, ys, mask, dim, shared, split_like)
    prev = v.score_
    is_dict = isinstance(prev, dict)
    phi = {name: [prev[name]] for name in prev} if is_dict else [prev]
    validators = [v]
    for p_i in perm.tolist():
        mask = mask | groups[p_i]
        v = fit_validator_(validator, Xs, ys, mask, dim, shared, split_like)
        validators.append(v)
        cur = v.score_
        if is_dict:
            for name in cur:
                phi[name].append(cur[name] - prev[name])
        else:
            phi.append(cur - prev)
        prev = cur
    sort_d = sort.to(X.device)
    if is_dict:
        phi = {name: torch.stack(vals, dim=0)[sort_d.to(vals[0].device)] for name, vals in phi.items()}
    else:
        phi = torch.stack(phi, dim=0)[sort_d.to(phi[0].device)]
    out = dict(score=phi, order=perm)
    out['validator'] = [validators[i] for i in sort.tolist()] if keep_validators else None
    return out


class Shapley:
    """Attributes: ``validator_`` (per permutation: validators in group order; ``None`` on other ranks),
    ``score_`` (n_permutations, n_groups, n_folds, ...), ``order_`` (n_permutations, n_groups - 1).
    ``score_.sum(1)`` equals the full model's score."""

    def __init__(self, validator: Validator, n_permutations: int = 10, n_jobs: Optional[int] = None,
                 verbose: Union[int, bool] = False, on_underspecified: str = 'raise', keep_validators: bool = True):
        self.validator = validator
        self.n_permutations = n_permutations
        self.n_jobs = n_jobs
        self.verbose = verbose
        self.on_underspecified = on_underspecified
        self.keep_validators = keep_validators      # the reference keeps all n_perm x n_groups validators in RAM
        self.validator_ = []
        self.score_ = []
        self.order_ = []

    def fit(self, X: torch.Tensor, y: torch.Tensor, groups=None, dim: Optional[int] = None) -> 'Shapley':
        if not (isinstance(X, torch.Tensor) and isinstance(y, torch.Tensor)):
            raise ValueError(f'Both `X` and `y` must be of type torch.Tensor, but got {type(X)} and {type(y)}.')
        dim, groups = prepare_groups_(X, groups, dim, self.on_underspecified)
        n_groups, P = groups.shape[0], self.n_permutations
        out_dev = X.device
        shard = _dist.sharding_active()
        # all RNG draws first, in the reference's order, on the input's device; rank 0 is authoritative
        if not shard or _dist.world()[0] == 0:
            draws = [draw_permutation_(n_groups, X.shape[0], X.device) for _ in range(P)]
            payload = [{k: v.cpu() for k, v in d.items()} for d in draws]
        else:
            payload = None
        if shard:
            payload = _dist.broadcast_object(payload, src=0)
        draws = [{k: v.to(X.device) for k, v in d.items()} for d in payload]
        mine = set(_dist.my_units(P)) if shard else set(range(P))
        Xd, yd = _lib.stage(X), _lib.stage(y)
        gd = groups.to(Xd.device)
        multi = [m.name for m in self.validator.metric] if (self.validator.metric is not None and len(self.validator.metric) > 1) else None
        self.validator_, local = [], {}
        with _dist.sharded_region():
            for p_i in range(P):
                if p_i not in mine:
                    self.validator_.append(None)
                    continue
                r = fit_permutation_(self.validator, Xd, yd, gd, dim, draws[p_i], keep_validators=self.keep_validators,
                                     orig=(X, y) if Xd.device != X.device else None)
                local[p_i] = r['score']
                self.validator_.append(_move_tensors(r['validator'], out_dev) if (r['validator'] is not None and Xd.device != out_dev) else r['validator'])
        if multi is not None:
            self.score_ = {}
            for name in multi:
                per = {i: s[name] for i, s in local.items()}
                per = _dist.gather_units(per, P) if shard else [per[i] for i in range(P)]
                self.score_[name] = torch.stack(per, dim=0).to(out_dev)
        else:
            per = _dist.gather_units(local, P) if shard else [local[i] for i in range(P)]
            self.score_ = torch.stack(per, dim=0).to(out_dev)
        self.order_ = torch.stack([d['perm'] for d in draws], dim=0).to(out_dev)
        return self
