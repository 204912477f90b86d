"""Repeated stratified k-fold splitter -- drop-in for ``mvpy.crossvalidation.RepeatedStratifiedKFold``
(mvpy/crossvalidation/repeatedstratifiedkfold.py:96-311): one shuffling stratified splitter split ``n_repeats``
times in a row.  Host-side integer work: no kernel involved."""
from __future__ import annotations

from typing import Optional, Union

import torch

from .stratifiedkfold import _StratifiedKFold_torch


class _RepeatedStratifiedKFold_torch:
    def __init__(self, n_splits: int = 5, n_repeats: int = 1, random_state: Optional[Union[int, torch.Generator]] = None):
        self.n_splits, self.n_repeats, self.random_state = n_splits, n_repeats, random_state
        self.kfold_ = _StratifiedKFold_torch(n_splits=n_splits, shuffle=True, random_state=random_state)

    def __repr__(self) -> str:
        return f'RepeatedStratifiedKFold(n_splits={self.n_splits}, n_repeats={self.n_repeats}, random_state={self.random_state})'

    def split(self, X: torch.Tensor, y: Optional[torch.Tensor] = None):
        for _ in range(self.n_repeats):
            yield from self.kfold_.split(X, y)


class RepeatedStratifiedKFold:
    """``split(X, y)`` yields ``n_repeats * n_splits`` (train, test) pairs; ``n_splits`` and ``n_repeats`` exposed."""

    def __init__(self, n_splits: int = 5, n_repeats: int = 1, random_state: Optional[Union[int, torch.Generator]] = None):
        self.n_splits, self.n_repeats, self.random_state = n_splits, n_repeats, random_state

    def __repr__(self) -> str:
        return f'RepeatedStratifiedKFold(n_splits={self.n_splits}, n_repeats={self.n_repeats}, random_state={self.random_state})'

    def split(self, X: torch.Tensor, y: Optional[torch.Tensor] = None):
        if isinstance(X, torch.Tensor) and isinstance(y, torch.Tensor):
            return self.to_torch().split(X, y)
        raise ValueError(f'X and y must be torch.Tensor for the B200 build, but got {type(X)} and {type(y)}.')

    def to_torch(self) -> _RepeatedStratifiedKFold_torch:
        return _RepeatedStratifiedKFold_torch(n_splits=self.n_splits, n_repeats=self.n_repeats, random_state=self.random_state)

    def to_numpy(self):
        raise NotImplementedError('The B200 build has no numpy backend.')
