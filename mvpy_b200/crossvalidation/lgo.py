"""Leave-groups-out splitter -- drop-in for ``mvpy.crossvalidation.LeaveGroupsOut``
(mvpy/crossvalidation/lgo.py:13-203).

The unique group labels are split by a repeated k-fold; a sample goes to the training (test) side when ALL of its
labels are training (test) labels, so samples with labels on both sides are in neither and the folds need not
partition the data.  Index tensors live on ``X.device`` and follow the reference bit for bit."""
from __future__ import annotations

from typing import Optional, Union

import torch

from .repeatedkfold import RepeatedKFold


class LeaveGroupsOut:
    def __init__(self, n_splits: int = 5, n_repeats: int = 1, random_state: Optional[Union[int, torch.Generator]] = None):
        self.n_splits, self.n_repeats, self.random_state = n_splits, n_repeats, random_state
        self.rkf_ = RepeatedKFold(n_splits=n_splits, n_repeats=n_repeats, random_state=random_state)

    def __repr__(self) -> str:
        return f'LeaveGroupsOut(n_splits={self.n_splits}, n_repeats={self.n_repeats}, random_state={self.random_state})'

    def split(self, X: torch.Tensor, y: Optional[torch.Tensor] = None, groups: Optional[torch.Tensor] = None):
        if groups is None:
            raise ValueError('`groups` must be specified when using LeaveGroupsOut.')
        if groups.shape[0] != X.shape[0]:
            raise ValueError(f'Number of samples must be equal across X and groups, but got {X.shape[0]} and {groups.shape[0]}, respectively.')
        if not (isinstance(X, torch.Tensor) and isinstance(groups, torch.Tensor) and (y is None or isinstance(y, torch.Tensor))):
            raise ValueError(f'`X`, `y` and `groups` must be torch.Tensor for the B200 build, but got {type(X)}, {type(y)} and {type(groups)}.')
        return self._split(groups)

    def _split(self, groups: torch.Tensor):
        labels = torch.unique(groups)
        label_dims = tuple(range(1, groups.ndim))
        for tr, te in self.rkf_.split(labels[:, None]):
            sides = []
            for chosen in (labels[tr], labels[te]):
                member = torch.isin(groups, chosen)
                if label_dims:
                    member = member.all(dim=label_dims)
                sides.append(torch.where(member)[0])
            yield sides[0], sides[1]

    def to_torch(self) -> 'LeaveGroupsOut':
        return self

    def to_numpy(self):
        raise NotImplementedError('The B200 build has no numpy backend.')
