"""Cross-validation driver -- drop-in for ``mvpy.crossvalidation.Validator``
(mvpy/crossvalidation/validator.py:20-680).

Per fold: clone the model, grant each metric the training-set baseline ``y_b = y[train].mean(reduce)``,
fit on ``X[train]``, score ``X[test]`` through ``metrics.score``, collect; scores are stacked to
(n_folds, ...).  Host tensors are moved to the GPU once, every fold runs there, and models / scores are
returned on the input's device.  Folds are independent, so under ``torch.distributed`` (one process per
GPU) they are sharded round-robin and the score shards are all-gathered (``_dist``); ``n_jobs`` is
accepted and ignored (there is no process pool).

Where the reference raises for evidently unintended reasons (SURVEY.md section 8c defects 3 and 4:
single-metric tuples, ``Validator.score`` with several metrics) this implementation returns what the
surrounding code was written to return."""
from __future__ import annotations

from copy import deepcopy
from typing import Any, Dict, List, Optional, Tuple, Union

import sklearn.base
import torch
from sklearn.pipeline import Pipeline

from .. import _dist, _lib, _streams
from ..metrics import Metric, score
from .kfold import KFold


def _move_tensors(obj: Any, device: torch.device, _seen=None) -> Any:
    """Return ``obj`` with every tensor reachable from it (estimator attributes, pipeline steps, lists,
    dicts, metric metadata) moved to ``device``; estimators are updated in place.  Views of one base tensor (the per-slice
    attributes of a batched ``Sliding`` fit are rows of a few stacked arrays) move with ONE copy of the base and stay views."""
    if _seen is None:
        _seen = {'ids': set(), 'bases': {}}
    if isinstance(obj, torch.Tensor):
        base = obj._base
        if base is None or obj.device == device or obj.requires_grad or base.numel() * base.element_size() > (256 << 20):
            return obj.to(device)
        key = (id(base), base.data_ptr())
        moved = _seen['bases'].get(key)
        if moved is None:
            moved = base.to(device)
            _seen['bases'][key] = moved
            _seen.setdefault('keep', []).append(base)           # keeps id(base) unique while the cache lives
        if moved.stride() != base.stride():
            return obj.to(device)
        return moved.as_strided(obj.size(), obj.stride(), obj.storage_offset() - base.storage_offset() + moved.storage_offset())
    if obj is None or isinstance(obj, (int, float, str, bool)) or id(obj) in _seen['ids']:
        return obj
    _seen['ids'].add(id(obj))
    if isinstance(obj, list):
        for i, o in enumerate(obj):
            obj[i] = _move_tensors(o, device, _seen)
        return obj
    if isinstance(obj, tuple):
        return tuple(_move_tensors(o, device, _seen) for o in obj)
    if isinstance(obj, dict):
        for k in list(obj):
            obj[k] = _move_tensors(obj[k], device, _seen)
        return obj
    if isinstance(obj, Pipeline):
        for _, step in obj.steps:
            _move_tensors(step, device, _seen)
        return obj
    if isinstance(obj, (sklearn.base.BaseEstimator, Metric)):
        stacked = getattr(obj, '_stacked', None)
        stack_all = getattr(obj, '_stack_all', None)
        if stack_all is not None and isinstance(getattr(obj, 'estimators_', None), list) and stack_all['coef'].device != device:
            # a batched Sliding fit going to another device (host inputs): the stacked arrays move, one copy each, and the
            # per-slice estimators are re-pointed at the moved rows -- instead of ~4 000 per-attribute moves per fold
            from .. import _batched
            obj._stack_all = _batched.move_stack(obj.estimators_, stack_all, device)
            obj._stacked = None                                  # (the compute-device copy for predict is dropped with the move)
            _seen['ids'].add(id(obj.estimators_))
            stacked = None
        for k, v in list(vars(obj).items()):
            if k == '_stack_all':
                continue
            if k == 'estimators_' and stacked is not None and isinstance(v, list) and v and stacked[0].device == device \
                    and getattr(v[0], 'coef_', None) is not None and v[0].coef_.device == device:
                continue        # a batched Sliding fit: hundreds of per-slice views of a few stacked arrays that already live on `device`
            if isinstance(v, torch.Tensor) and k.startswith('_'):
                setattr(obj, k, None)              # private device scratch (cached Gram) is dropped
            elif isinstance(v, (torch.Tensor, list, tuple, dict, sklearn.base.BaseEstimator, Pipeline, Metric)):
                setattr(obj, k, _move_tensors(v, device, _seen))
    return obj


def _resolve_metric(model, metric):
    """Default metric (validator.py:59-66): the estimator's (or the pipeline's last step's) ``metric_``."""
    if metric is not None:
        return metric
    if isinstance(model, Pipeline):
        return getattr(model[-1], 'metric_', None)
    return getattr(model, 'metric_', None)


def _lanes_ok(model) -> bool:
    """Folds may run on concurrent stream lanes when the fit is one of the large launch chains that profit (``TimeDelayed``
    / ``RidgeCV`` behind per-feature preprocessing); estimators that manage streams or CUDA graphs themselves (``Sliding``)
    and user-supplied models keep the caller's stream."""
    last = model[-1] if isinstance(model, Pipeline) else model
    return bool(getattr(last, '_mvb_fold_lanes', False))


def fold_key(train: torch.Tensor):
    """Host-side identity of a training set (key of the predictor-set drivers' per-fold statistics cache); computed by the
    driver's own thread before the fold is enqueued, so that no fit has to read its indices back from the device."""
    return hash(train.detach().cpu().numpy().tobytes())


def fit_model_(model, train: torch.Tensor, test: torch.Tensor, X: torch.Tensor, y: Optional[torch.Tensor] = None,
               metric: Optional[Union[Metric, Tuple[Metric]]] = None, train_key=None) -> Dict:
    """One fold (validator.py:20-115)."""
    model = model.clone() if hasattr(model, 'clone') else deepcopy(model)
    metric = deepcopy(_resolve_metric(model, metric))
    many = isinstance(metric, tuple)
    if y is not None and metric is not None:
        for m in (metric if many else (metric,)):
            m.grant({'y_b': y[train].mean(m.reduce, keepdim=True)})
    _lib.fold_ctx.train, _lib.fold_ctx.model, _lib.fold_ctx.train_key = train, model, train_key    # lets TimeDelayed share per-fold statistics across predictor sets
    try:
        model.fit(*((X[train], y[train]) if y is not None else (X[train],)))
    finally:
        _lib.fold_ctx.train = _lib.fold_ctx.model = _lib.fold_ctx.train_key = None
    args = (X[test], y[test]) if y is not None else (X[test],)
    score_i = score(model, metric, *args) if metric is not None else model.score(*args)
    if many and len(metric) > 1:
        score_ = {m.name: score_i[m.name] for m in metric}
    else:
        score_ = score_i
    return dict(model=model, score_=score_, test=test, metric=metric)


class Validator(sklearn.base.BaseEstimator):
    """Attributes after ``fit``: ``model_`` (fitted clone per fold; ``None`` for folds fitted on another
    rank), ``score_`` ((n_folds, ...) tensor, or dict of them for several metrics), ``test_``, ``metric_``,
    ``cv_``, ``cv_n_``."""

    def __init__(self, model, cv: Union[int, Any] = 5, metric: Optional[Union[Metric, Tuple[Metric]]] = None,
                 n_jobs: Optional[int] = None, verbose: Union[int, bool] = False):
        self.model = model
        self.cv = cv
        self.metric = metric
        self.n_jobs = n_jobs
        self.verbose = verbose
        if self.metric is None and not hasattr(model, 'score'):
            raise ValueError('If no `metric` is specified, `model` must expose a `score()`-method, but none was found.')
        if isinstance(self.metric, Metric):
            self.metric = (self.metric,)
        self.cv_ = KFold(n_splits=self.cv) if isinstance(cv, int) else cv
        if not hasattr(self.cv_, 'split'):
            raise ValueError('Expected `cv` to be either an integer or expose a `split()` method.')
        if not hasattr(self.cv_, 'n_splits'):
            raise ValueError('Expceted `cv` to expose `n_splits`.')
        self.cv_n_ = self.cv_.n_splits
        if hasattr(self.cv_, 'n_repeats'):
            self.cv_n_ = self.cv_n_ * self.cv_.n_repeats
        self.model_ = []
        self.score_ = []
        self.test_ = []
        self.metric_ = []

    def _multi(self) -> bool:
        return self.metric is not None and len(self.metric) > 1

    def fit(self, X: torch.Tensor, y: Optional[torch.Tensor] = None, groups: Optional[torch.Tensor] = None,
            _split_like: Optional[Tuple] = None) -> 'Validator':
        self.model_, self.test_, self.metric_ = [], [], []
        out_dev = X.device
        # indices are generated on the input's device exactly like the reference (bit-exact folds); a driver that
        # has already staged host data on the GPU hands over the original tensors (`_split_like`) so that shuffling
        # splitters still draw from the generator of the device the user's data lives on
        split_args = _split_like if _split_like is not None else ((X, y) if y is not None else (X,))
        split_kwargs = dict(groups=groups) if groups is not None else dict()
        folds = list(self.cv_.split(*split_args, **split_kwargs))
        Xd, yd = _lib.stage(X), _lib.stage(y)
        shard = _dist.sharding_active()
        mine = set(_dist.my_units(len(folds))) if shard else set(range(len(folds)))
        local_scores: Dict[int, Any] = {}
        todo = [k for k in range(len(folds)) if k in mine]

        keys = {k: fold_key(folds[k][0]) for k in todo} if _lib.fold_ctx.parent is not None else {}

        def one(k: int):
            train, test = folds[k]
            r = fit_model_(self.model, train.to(Xd.device), test.to(Xd.device), Xd, y=yd, metric=self.metric, train_key=keys.get(k))
            # host inputs: the fitted model goes back to the input's device here, i.e. on this fold's lane, so that its
            # device-to-host copies overlap the other folds' kernels
            r['model'], r['metric'] = _move_tensors(r['model'], out_dev), _move_tensors(r['metric'], out_dev)
            return r
        with _dist.sharded_region():           # process-wide depth counter: held by this thread while the lanes work
            # large fits leave SMs idle inside their launch chains: folds go to a few stream lanes (fold k always on lane k mod n)
            if Xd.is_cuda and _lanes_ok(self.model):
                done = _streams.run_on_lanes([(lambda k=k: one(k)) for k in todo], Xd.device, lane_keys=todo)
            else:
                done = [one(k) for k in todo]
        fitted = dict(zip(todo, done))
        for k, (train, test) in enumerate(folds):
            self.test_.append(test)
            r = fitted.get(k)
            if r is None:
                self.model_.append(None)
                self.metric_.append(None)
                continue
            local_scores[k] = r['score_']
            self.model_.append(r['model'])
            self.metric_.append(r['metric'])
        if self._multi():
            self.score_ = {}
            for m in self.metric:
                per = {k: v[m.name] for k, v in local_scores.items()}
                per = _dist.gather_units(per, len(folds)) if shard else [per[k] for k in range(len(folds))]
                self.score_[m.name] = torch.stack(per, dim=0).to(out_dev)
        else:
            per = _dist.gather_units(local_scores, len(folds)) if shard else [local_scores[k] for k in range(len(folds))]
            self.score_ = torch.stack(per, dim=0).to(out_dev)
        return self

    # -- convenience accessors (validator.py:402-640) ----------------------------------------------
    def _fitted(self):
        if len(self.model_) == 0:
            raise ValueError('Validator has not been fitted yet.')
        return [m for m in self.model_ if m is not None]

    def _targets(self, from_step, what: str, name: str):
        models = self._fitted()
        if isinstance(models[0], Pipeline):
            if from_step is not None:
                if not hasattr(models[0][from_step], name):
                    raise ValueError(f'Estimator at position {from_step} in Pipeline does not expose {what}.')
                return [m[from_step] for m in models]
            if not hasattr(models[0], name):
                raise ValueError(f'Pipeline does not expose {what}.')
            return models
        if not hasattr(models[0], name):
            raise ValueError(f'Model does not expose {what}.')
        return models

    @staticmethod
    def _stack(out: List[Any]):
        return torch.stack(out, dim=0) if isinstance(out[0], torch.Tensor) else out

    def call(self, method: str, *args: Any, from_step: Optional[int] = None, **kwargs: Any):
        return self._stack([getattr(t, method)(*args, **kwargs) for t in self._targets(from_step, f'{method}()', method)])

    def collect(self, attr: str, from_step: Optional[int] = None):
        return self._stack([getattr(t, attr) for t in self._targets(from_step, f'attribute {attr}', attr)])

    def transform(self, X):
        return self.call('transform', X)

    def decision_function(self, X):
        return self.call('decision_function', X)

    def predict(self, X):
        return self.call('predict', X)

    def predict_proba(self, X):
        return self.call('predict_proba', X)

    def score(self, X: torch.Tensor, y: torch.Tensor):
        """Score new data with every fold's model and its granted metric (training baselines kept)."""
        args = (X, y) if y is not None else (X,)
        models = self._fitted()
        metrics_ = [m for m in self.metric_ if m is not None]
        scores_ = [score(mod, met, *args) if met is not None else mod.score(*args) for mod, met in zip(models, metrics_)]
        if isinstance(scores_[0], dict):
            return {name: torch.stack([s[name] for s in scores_], dim=0) for name in scores_[0]}
        return torch.stack(scores_, dim=0)

    def clone(self) -> 'Validator':
        return Validator(self.model, cv=self.cv, metric=self.metric, n_jobs=self.n_jobs, verbose=self.verbose)
