"""One-hot label encoding for multi-class, multi-feature targets -- drop-in for ``mvpy.preprocessing.LabelBinariser``
(mvpy/preprocessing/labelbinariser.py:228-700).

Every label feature (column of ``y``) gets one output column per class (sorted label order), features side by side;
``pos_label`` marks the sample's class, ``neg_label`` the rest.  Index work on the labels' device (no kernel of the
library is involved); the output dtype follows ``neg_label`` like the reference's ``torch.full``."""
from __future__ import annotations

from typing import Any, List, Optional

import sklearn.base
import torch


class _LabelBinariser_torch(sklearn.base.BaseEstimator):
    """Attributes after ``fit``: ``n_features_, n_classes_`` (per feature), ``labels_`` (sorted labels per feature),
    ``classes_`` (0..k-1 per feature), ``N_`` (total columns), ``C_`` (first column of every feature), ``map_L_to_C_``."""

    def __init__(self, neg_label: int = 0, pos_label: int = 1):
        if neg_label >= pos_label:
            raise ValueError('Negative label must be smaller than positive label.')
        self.neg_label = neg_label
        self.pos_label = pos_label
        self.n_features_ = None
        self.n_classes_ = None
        self.labels_: Optional[List[torch.Tensor]] = None
        self.classes_ = None
        self.N_ = None
        self.C_ = None
        self.map_L_to_C_ = None

    def fit(self, y: torch.Tensor, *args: Any):
        if y.ndim == 1:
            y = y[:, None]
        self.n_features_ = y.shape[1]
        self.labels_ = [torch.unique(y[:, i]) for i in range(self.n_features_)]
        self.n_classes_ = [int(l.shape[0]) for l in self.labels_]
        self.classes_ = [torch.arange(k) for k in self.n_classes_]
        counts = torch.tensor(self.n_classes_, dtype=y.dtype, device=y.device)
        self.N_ = counts.sum()
        self.C_ = torch.cat((counts.new_zeros(1), counts.cumsum(0)[:-1]))
        self.map_L_to_C_ = [{label: cls for cls, label in zip(self.classes_[i], self.labels_[i])} for i in range(self.n_features_)]
        return self

    def _check_fitted(self, what: str) -> None:
        if self.labels_ is None or self.classes_ is None or self.map_L_to_C_ is None:
            raise ValueError(f'LabelBinariser must be fit before calling {what}.')

    def transform(self, y: torch.Tensor, *args: Any) -> torch.Tensor:
        if y.ndim == 1:
            y = y[:, None]
        self._check_fitted('transform')
        if y.shape[1] != self.n_features_:
            raise ValueError(f'LabelBinariser expected {self.n_features_} features, but got {y.shape[1]}.')
        out = torch.full((y.shape[0], int(self.N_.item())), self.neg_label, device=y.device)
        start = 0
        for i, labels in enumerate(self.labels_):
            labels = labels.to(y.device)
            hit = y[:, i, None] == labels[None, :]                          # (n, k) membership of every sample
            if not hit.any(1).all():
                raise ValueError('Unknown labels in `y`.')
            block = out[:, start:start + labels.shape[0]]
            block[hit] = self.pos_label
            start += labels.shape[0]
        return out

    def fit_transform(self, y: torch.Tensor, *args: Any) -> torch.Tensor:
        return self.fit(y, *args).transform(y, *args)

    def inverse_transform(self, y: torch.Tensor, *args: Any) -> torch.Tensor:
        if y.ndim == 1:
            y = y[:, None]
        self._check_fitted('inverse_transform')
        if y.shape[1] != self.N_:
            raise ValueError(f'LabelBinariser expected {self.N_} subfeatures, but got {y.shape[1]}.')
        cols, start = [], 0
        for labels in self.labels_:
            block = y[:, start:start + labels.shape[0]]
            cols.append(labels.to(y.device)[torch.argmax((block == self.pos_label).long(), dim=1)])
            start += labels.shape[0]
        return torch.stack(cols).T

    def clone(self):
        return _LabelBinariser_torch(neg_label=self.neg_label, pos_label=self.pos_label)

    def copy(self):
        return self.clone()


class LabelBinariser(sklearn.base.BaseEstimator):
    """Front end (labelbinariser.py:446-700): every call builds a torch binariser for torch input."""

    def __init__(self, neg_label: int = 0, pos_label: int = 1):
        self.neg_label = neg_label
        self.pos_label = pos_label

    def _make(self, y):
        if isinstance(y, torch.Tensor):
            return _LabelBinariser_torch(neg_label=self.neg_label, pos_label=self.pos_label)
        raise TypeError(f'`y` must be torch.Tensor (the B200 build has no numpy backend), but got {type(y)}.')

    def fit(self, y, *args):
        return self._make(y).fit(y, *args)

    def transform(self, y, *args):
        return self._make(y).transform(y, *args)

    def inverse_transform(self, y, *args):
        return self._make(y).inverse_transform(y, *args)

    def fit_transform(self, y, *args):
        return self._make(y).fit_transform(y, *args)

    def to_torch(self):
        return _LabelBinariser_torch(neg_label=self.neg_label, pos_label=self.pos_label)

    def to_numpy(self):
        raise NotImplementedError('The B200 build has no numpy backend.')

    def clone(self):
        return LabelBinariser(neg_label=self.neg_label, pos_label=self.pos_label)

    def copy(self):
        return self.clone()
