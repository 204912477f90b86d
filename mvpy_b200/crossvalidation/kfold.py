"""K-fold splitter -- drop-in for ``mvpy.crossvalidation.KFold`` (mvpy/crossvalidation/kfold.py:170-446).

Index generation is host-side integer work and must be bit-exact with the reference:
contiguous test blocks of n//k samples (+1 for the first n%k folds) over ``arange(n)`` or a seeded
``torch.randperm`` drawn on ``X.device`` exactly as the reference does."""
from __future__ import annotations

from collections.abc import Generator
from typing import Optional, Union

import torch


class _KFold_torch:
    def __init__(self, n_splits: int = 5, shuffle: bool = False, random_state: Optional[Union[int, torch.Generator]] = None):
        self.n_splits = n_splits
        self.shuffle = shuffle
        self.random_state = random_state
        if isinstance(random_state, int) or random_state is None:
            self.rng_ = torch.Generator()
            if random_state is not None:
                self.rng_.manual_seed(random_state)
        else:
            self.rng_ = random_state

    def __repr__(self) -> str:
        return f'KFold(n_splits={self.n_splits}, random_state={self.random_state}, shuffle={self.shuffle})'

    def split(self, X: torch.Tensor, y: Optional[torch.Tensor] = None) -> Generator[tuple[torch.Tensor, torch.Tensor], None, None]:
        n = X.shape[0]
        if n < self.n_splits:
            raise ValueError(f'n_samples must be greater than n_splits, but got {n} and {self.n_splits}.')
        if self.shuffle:
            if self.rng_.device != X.device:          # generator follows the data's device (kfold.py:212-218)
                self.rng_ = torch.Generator(device=X.device)
                if self.random_state is not None:
                    self.rng_.manual_seed(self.random_state)
            order = torch.randperm(n, generator=self.rng_, device=X.device).long()
        else:
            order = torch.arange(n, device=X.device).long()
        lo = 0
        for k in range(self.n_splits):
            hi = lo + n // self.n_splits + (1 if k < n % self.n_splits else 0)
            yield torch.cat((order[:lo], order[hi:])), order[lo:hi]
            lo = hi


class KFold:
    """``split(X[, y])`` yields (train, test) long index tensors on ``X.device``; ``n_splits`` exposed."""

    def __init__(self, n_splits: int = 5, shuffle: bool = False, random_state: Optional[Union[int, torch.Generator]] = None):
        self.n_splits = n_splits
        self.shuffle = shuffle
        self.random_state = random_state

    def __repr__(self) -> str:
        return f'KFold(n_splits={self.n_splits}, random_state={self.random_state}, shuffle={self.shuffle})'

    def split(self, X: torch.Tensor, y: Optional[torch.Tensor] = None):
        if isinstance(X, torch.Tensor) and (y is None or isinstance(y, torch.Tensor)):
            return self.to_torch().split(X, y)
        raise ValueError(f'X (and y) must be torch.Tensor for the B200 build, but got {type(X)} and {type(y)}.')

    def to_torch(self) -> _KFold_torch:
        return _KFold_torch(n_splits=self.n_splits, shuffle=self.shuffle, random_state=self.random_state)

    def to_numpy(self):
        raise NotImplementedError('The B200 build has no numpy backend.')
