"""Repeated k-fold splitter -- drop-in for ``mvpy.crossvalidation.RepeatedKFold``
(mvpy/crossvalidation/repeatedkfold.py:14-310).

One shuffling k-fold splitter is created per object and split ``n_repeats`` times in a row, so the repeats
consume one random stream (seeded by ``random_state``); indices are long tensors on ``X.device`` and bit-exact
with the reference.  Host-side integer work: no kernel involved."""
from __future__ import annotations

from typing import Optional, Union

import torch

from .kfold import _KFold_torch


class _RepeatedKFold_torch:
    def __init__(self, n_splits: int = 5, n_repeats: int = 1, random_state: Optional[Union[int, torch.Generator]] = None):
        self.n_splits, self.n_repeats, self.random_state = n_splits, n_repeats, random_state
        self.kfold_ = _KFold_torch(n_splits=n_splits, shuffle=True, random_state=random_state)

    def __repr__(self) -> str:
        return f'RepeatedKFold(n_splits={self.n_splits}, n_repeats={self.n_repeats}, random_state={self.random_state})'

    def split(self, X: torch.Tensor, y: Optional[torch.Tensor] = None):
        for _ in range(self.n_repeats):
            yield from self.kfold_.split(X, y)


class RepeatedKFold:
    """``split(X[, y])`` yields ``n_repeats * n_splits`` (train, test) pairs; ``n_splits`` and ``n_repeats`` exposed
    (``Validator`` multiplies them, validator.py:303-313).  Like the reference, every ``split`` call on this
    front-end object starts a fresh splitter (and so a fresh seeded stream); use ``to_torch()`` to keep one."""

    def __init__(self, n_splits: int = 5, n_repeats: int = 1, random_state: Optional[Union[int, torch.Generator]] = None):
        self.n_splits, self.n_repeats, self.random_state = n_splits, n_repeats, random_state

    def __repr__(self) -> str:
        return f'RepeatedKFold(n_splits={self.n_splits}, n_repeats={self.n_repeats}, random_state={self.random_state})'

    def split(self, X: torch.Tensor, y: Optional[torch.Tensor] = None):
        if isinstance(X, torch.Tensor) and (y is None or isinstance(y, torch.Tensor)):
            return self.to_torch().split(X, y)
        raise ValueError(f'X (and y) must be torch.Tensor for the B200 build, but got {type(X)} and {type(y)}.')

    def to_torch(self) -> _RepeatedKFold_torch:
        return _RepeatedKFold_torch(n_splits=self.n_splits, n_repeats=self.n_repeats, random_state=self.random_state)

    def to_numpy(self):
        raise NotImplementedError('The B200 build has no numpy backend.')
