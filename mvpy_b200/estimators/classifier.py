"""One-versus-rest / one-versus-one wrapper around a classifier type -- drop-in for ``mvpy.estimators.Classifier``
(mvpy/estimators/classifier.py:655-1308).

``OvR`` hands the labels to one estimator (which binarises them itself and takes the largest decision value per label
feature).  ``OvO`` fits one estimator per pair of classes of every label feature on the samples of those two classes
and predicts by majority vote.  Host-side orchestration only: every fit / prediction is the wrapped estimator's CUDA
path.  ``predict_proba`` for ``OvO`` couples the pairwise probabilities (Wu-Lin, classifier.py:573-653, 913-1000) on the
``mvb_pairwise_coupling`` kernel."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple, Union

import sklearn.base
import torch

from .. import _lib, metrics
from ..preprocessing.labelbinariser import _LabelBinariser_torch


class _Classifier_torch(sklearn.base.BaseEstimator):
    """Attributes: ``estimator`` (a class), ``method, arguments, kwarguments``; after ``fit``: ``estimators_`` (one
    instance, or the list of pairwise instances), ``binariser_, coef_, intercept_, pattern_`` (stacked over the pairwise
    classifiers for OvO), ``offsets_``; ``metric_`` = accuracy."""

    def __init__(self, estimator, method: str = 'OvR', arguments: List[Any] = [], kwarguments: Dict[Any, Any] = dict()):
        self.estimator = estimator
        self.method = method
        self.arguments = arguments
        self.kwarguments = kwarguments
        if method not in ['OvR', 'OvO']:
            raise ValueError(f'Method `{method}` unknown. Must be [\'OvR\', \'OvO\'].')
        self.estimators_ = None
        self.binariser_ = _LabelBinariser_torch()
        self.coef_ = None
        self.intercept_ = None
        self.pattern_ = None
        self.offsets_ = None
        self.metric_ = metrics.accuracy

    def _pairs(self):
        """(feature, class j, class k) with j > k, in the reference's loop order (classifier.py:777-781)."""
        for i in range(self.binariser_.n_features_):
            n = self.binariser_.n_classes_[i]
            for j in range(n):
                for k in range(n):
                    if j > k:
                        yield i, j, k

    def fit(self, X: torch.Tensor, y: torch.Tensor):
        if len(y.shape) == 1:
            y = y[:, None]
        if self.method == 'OvR':
            self.estimators_ = self.estimator(*self.arguments, **self.kwarguments).fit(X, y)
            for attr in ('coef_', 'intercept_', 'pattern_'):
                if hasattr(self.estimators_, attr):
                    setattr(self, attr, getattr(self.estimators_, attr))
            return self
        self.binariser_.fit(y)
        self.estimators_ = []
        offsets, offset = [], 0
        for i in range(self.binariser_.n_features_):
            offsets.append(offset)
        collected = {'coef_': [], 'intercept_': [], 'pattern_': []}
        for i, j, k in self._pairs():
            labels = self.binariser_.labels_[i]
            both = (y[:, i] == labels[j]) | (y[:, i] == labels[k])
            est = self.estimator(*self.arguments, **self.kwarguments).fit(X[both], y[both, i])
            self.estimators_.append(est)
            for attr, bucket in collected.items():
                if hasattr(est, attr):
                    bucket.append(getattr(est, attr))
        self.coef_ = torch.stack(collected['coef_']) if collected['coef_'] else torch.tensor([], dtype=X.dtype, device=X.device)
        self.intercept_ = torch.stack(collected['intercept_'])
        self.pattern_ = torch.stack(collected['pattern_']) if collected['pattern_'] else []
        self.offsets_ = torch.tensor(offsets)
        return self

    def _check(self, what: str) -> None:
        if self.estimators_ is None:
            raise ValueError(f'Classifier must be fit before calling {what}.')

    def decision_function(self, X: torch.Tensor) -> torch.Tensor:
        self._check('decision function')
        if self.method == 'OvR':
            return self.estimators_.decision_function(X)
        return torch.stack([est.decision_function(X) for est in self.estimators_])

    def predict(self, X: torch.Tensor) -> torch.Tensor:
        self._check('predict')
        if self.method == 'OvR':
            return self.estimators_.predict(X)
        pair_of = {}
        for idx, (i, j, k) in enumerate(self._pairs()):
            pair_of.setdefault(i, []).append(idx)
        columns = []
        for i in range(self.binariser_.n_features_):
            votes = torch.stack([self.estimators_[idx].predict(X).squeeze() for idx in pair_of[i]])      # (pairs, n)
            labels = self.binariser_.labels_[i].to(votes.device)
            counts = (votes[:, :, None] == labels[None, None, :]).to(X.dtype).sum(0)                     # (n, classes)
            columns.append(labels[torch.argmax(counts, dim=1)])
        return torch.stack(columns).T

    def predict_proba(self, X: torch.Tensor) -> torch.Tensor:
        self._check('predict')
        if self.method == 'OvR':
            return self.estimators_.predict_proba(X)
        out_dev = X.device
        p_ind = 1 / (1 + torch.exp(-_lib.to_device(self.decision_function(X))))              # (pairs, n, 2)
        ijk, p = 0, []
        for i in range(self.binariser_.n_features_):
            K = int(self.binariser_.n_classes_[i])
            j_i, k_i = torch.tril_indices(K, K, offset=-1, device=p_ind.device)             # pairwise estimators: lower-triangle order
            span = torch.arange(ijk, ijk + j_i.shape[0], device=p_ind.device)
            p_i = torch.zeros((X.shape[0], K, K), dtype=p_ind.dtype, device=p_ind.device)
            p_i[:, k_i, j_i] = p_ind[span, :, 0].t()
            p_i[:, j_i, k_i] = p_ind[span, :, 1].t()
            ijk += j_i.shape[0]
            eps = 5e-3 / K
            p.append(_lib.pairwise_coupling(p_i, max(100, K), eps, eps))
        return torch.cat(p, dim=-1).to(out_dev)

    def score(self, X: torch.Tensor, y: torch.Tensor, metric: Optional[Union[metrics.Metric, Tuple[metrics.Metric]]] = None
              ) -> Union[torch.Tensor, Dict[str, torch.Tensor]]:
        if metric is None:
            metric = self.metric_
        return metrics.score(self, metric, X, y)

    def clone(self):
        return _Classifier_torch(estimator=self.estimator, method=self.method, arguments=self.arguments, kwarguments=self.kwarguments)


class Classifier(sklearn.base.BaseEstimator):
    def __new__(cls, estimator, method: str = 'OvR', arguments: List[Any] = [], kwarguments: Dict[Any, Any] = dict()):
        return _Classifier_torch(estimator, method=method, arguments=arguments, kwarguments=kwarguments)
