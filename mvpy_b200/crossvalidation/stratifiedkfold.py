"""Stratified k-fold splitter -- drop-in for ``mvpy.crossvalidation.StratifiedKFold``
(mvpy/crossvalidation/stratifiedkfold.py:158-470).

Samples are grouped by their joint label over all target features; every group is dealt over the folds in
round-robin proportion so each fold keeps the label mix.  Host-side integer work (no kernel); indices are long
tensors on ``X.device``, bit-exact with the reference, including two of its quirks: the joint label is a float32
weighted sum of the one-hot columns (stratifiedkfold.py:262-265), and shuffling draws from torch's *global*
generator of ``X.device`` -- the object's own ``rng_`` is re-seeded but never passed on (:292-300)."""
from __future__ import annotations

from typing import Optional, Union

import torch


def _one_hot_columns(y: torch.Tensor) -> torch.Tensor:
    """(n, sum of classes per feature) 0/1 matrix, features in order, each feature's classes in sorted label
    order -- the layout ``LabelBinariser(0, 1).fit_transform`` produces (preprocessing/labelbinariser.py:288-373)."""
    cols = []
    for i in range(y.shape[1]):
        labels, inverse = torch.unique(y[:, i], return_inverse=True)
        cols.append(torch.nn.functional.one_hot(inverse, labels.shape[0]))
    return torch.cat(cols, dim=1)


class _StratifiedKFold_torch:
    def __init__(self, n_splits: int = 5, shuffle: bool = False, random_state: Optional[Union[int, torch.Generator]] = None):
        self.n_splits, self.shuffle, self.random_state = n_splits, shuffle, random_state
        if isinstance(random_state, int) or random_state is None:
            self.rng_ = torch.Generator()
            if random_state is not None:
                self.rng_.manual_seed(random_state)
        else:
            self.rng_ = random_state

    def __repr__(self) -> str:
        return f'StratifiedKFold(n_splits={self.n_splits}, random_state={self.random_state}, shuffle={self.shuffle})'

    def _groups(self, y: torch.Tensor) -> torch.Tensor:
        """Group id per sample, numbered by first appearance."""
        if y.ndim == 1:
            y = y[:, None]
        if y.ndim > 2:                                    # (n, ..., features, time): first entry of every other dim
            pick = [0] * y.ndim
            pick[0] = pick[-2] = slice(None)
            y = y[tuple(pick)]
        hot = _one_hot_columns(y).float()
        weight = (2 ** torch.arange(hot.shape[1] - 1, -1, -1, device=y.device)).float()
        code = (hot * weight[None, :]).sum(1).long()
        values, inverse = torch.unique(code, return_inverse=True)
        first = torch.full((values.shape[0],), code.shape[0], dtype=torch.long, device=y.device)
        first.scatter_reduce_(0, inverse, torch.arange(code.shape[0], device=y.device), reduce='amin')
        return torch.argsort(torch.argsort(first))[inverse]

    def split(self, X: torch.Tensor, y: Optional[torch.Tensor] = None):
        n = X.shape[0]
        if n < self.n_splits:
            raise ValueError(f'n_samples must be greater than n_splits, but got {n} and {self.n_splits}.')
        group = self._groups(y)
        counts = torch.bincount(group)
        n_groups = counts.shape[0]
        if torch.all(self.n_splits > counts):
            raise ValueError('`n_splits` must be smaller or equal to the number of members in each class.')
        ordered, _ = torch.sort(group)
        share = torch.stack([torch.bincount(ordered[i::self.n_splits], minlength=n_groups) for i in range(self.n_splits)])
        fold_ids = torch.arange(self.n_splits, device=X.device)
        test_fold = torch.zeros(n, dtype=torch.long, device=X.device)
        for g in range(n_groups):
            deal = torch.repeat_interleave(fold_ids, share[:, g])
            if self.shuffle:
                if self.rng_.device != X.device:
                    self.rng_ = torch.Generator(device=X.device)
                    if self.random_state is not None:
                        self.rng_.manual_seed(self.random_state)
                deal = deal[torch.randperm(deal.shape[0], device=X.device).long()]     # global generator, see module doc
            test_fold[group == g] = deal
        index = torch.arange(n, device=X.device)
        for i in range(self.n_splits):
            held = test_fold == i
            yield index[~held], index[held]


class StratifiedKFold:
    """``split(X, y)`` yields (train, test) long index tensors on ``X.device``; ``n_splits`` exposed."""

    def __init__(self, n_splits: int = 5, shuffle: bool = False, random_state: Optional[Union[int, torch.Generator]] = None):
        self.n_splits, self.shuffle, self.random_state = n_splits, shuffle, random_state

    def __repr__(self) -> str:
        return f'StratifiedKFold(n_splits={self.n_splits}, random_state={self.random_state}, shuffle={self.shuffle})'

    def split(self, X: torch.Tensor, y: Optional[torch.Tensor] = None):
        if isinstance(X, torch.Tensor) and isinstance(y, torch.Tensor):
            return self.to_torch().split(X, y)
        raise ValueError(f'X and y must be torch.Tensor for the B200 build, but got {type(X)} and {type(y)}.')

    def to_torch(self) -> _StratifiedKFold_torch:
        return _StratifiedKFold_torch(n_splits=self.n_splits, shuffle=self.shuffle, random_state=self.random_state)

    def to_numpy(self):
        raise NotImplementedError('The B200 build has no numpy backend.')
