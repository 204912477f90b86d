"""Sliding estimator -- drop-in for ``mvpy.estimators.Sliding`` (mvpy/estimators/sliding.py:520-1227).

One clone of the wrapped estimator (or one call of the wrapped callable) per slice of ``dims[0]``;
nested dims recurse.  Each per-slice fit runs on the GPU through the wrapped estimator's own CUDA
path; ``n_jobs`` is accepted for API compatibility (the slices are independent launches on one
stream, there is no process pool)."""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, Optional, Tuple, Union

import numpy as np
import sklearn.base
import torch

from .. import _batched, _graphs, metrics


class _Sliding_torch(sklearn.base.BaseEstimator):
    """Attributes: ``estimator, dims, n_jobs, top, verbose, estimators_``."""

    def __init__(self, estimator: Union[Callable, sklearn.base.BaseEstimator], dims: tuple = (-1,), n_jobs: Union[int, None] = None,
                 top: bool = True, verbose: bool = True):
        self.estimator = estimator
        self.dims = dims
        self.n_jobs = n_jobs
        self.verbose = verbose
        self.top = top
        self.estimators_ = None
        self._stacked = None
        self._stack_all = None

    # (Folds of a batched Sliding fit do NOT take the Validator's stream lanes: measured at configs[4], 3 lanes 96 k fits/s against
    # 118 k serial -- the per-fold host work is Python under the GIL and the batched kernels already fill the GPU.)

    # -- helpers --------------------------------------------------------------------------------
    def _moved(self, X: torch.Tensor, y: Optional[torch.Tensor]):
        Xm = torch.moveaxis(X, self.dims[0], -1)
        ym = torch.moveaxis(y, self.dims[0], -1) if y is not None else None
        eff = tuple((torch.arange(len(Xm.shape))[list(self.dims)] - 1).tolist())
        return Xm, ym, eff

    @staticmethod
    def _pairs(Xm: torch.Tensor, ym: torch.Tensor):
        """slice indices; a singleton sliding dim on one side is broadcast (sliding.py:606-611)."""
        nx, ny = Xm.shape[-1], ym.shape[-1]
        xi = range(nx) if nx > 1 else [0] * ny
        yi = range(ny) if ny > 1 else [0] * nx
        return list(zip(xi, yi))

    def fit(self, X: torch.Tensor, y: torch.Tensor, *args: Any):
        if (isinstance(X, torch.Tensor) == False) | (isinstance(y, torch.Tensor) == False):
            raise ValueError(f'X and y must be of type torch.Tensor but got {type(X)} and {type(y)}.')
        if len(X.shape) != len(y.shape):
            raise ValueError(f'X and y must have the same number of dimensions for alignment, but got {X.shape} and {y.shape}.')
        if isinstance(self.estimator, sklearn.base.BaseEstimator):
            Xm, ym, eff = self._moved(X, y)
            proto = _Sliding_torch(estimator=self.estimator, dims=eff[1:], n_jobs=None, top=False) if len(self.dims) > 1 \
                else self.estimator.clone()
            pairs = self._pairs(Xm, ym)
            # many small identical fits: one batched call for the whole stack (mvb_dense_ridge_batched); shapes it does not
            # take are captured once as a CUDA graph and replayed per slice on concurrent streams
            fitted, self._stacked, self._stack_all = None, None, None
            if len(self.dims) == 1 and not args:
                got = _batched.fit_slices(proto, Xm, ym, pairs)
                if got is not None:
                    fitted, self._stacked, self._stack_all = got
                else:
                    fitted = _graphs.fit_slices(proto, Xm, ym, pairs)
            self.estimators_ = fitted if fitted is not None else [proto.clone().fit(Xm[..., i], ym[..., j], *args) for i, j in pairs]
        else:
            self.estimators_ = []
        return self

    def _apply(self, method: str, X: torch.Tensor, y: Optional[torch.Tensor], *args) -> torch.Tensor:
        if isinstance(X, torch.Tensor) == False:
            raise ValueError(f'X must be of type torch.Tensor but got {type(X)}.')
        if (self.estimators_ is None) & (callable(self.estimator) == False):
            raise ValueError('Estimators not fitted yet.')
        if isinstance(self.estimator, sklearn.base.BaseEstimator):
            Xm = torch.moveaxis(X, self.dims[0], -1)
            if method == 'predict' and not args and len(self.dims) == 1:
                yh = _batched.predict_slices(self.estimators_, Xm, getattr(self, '_stacked', None))     # (N, F, slices) in one launch
                if yh is not None:
                    return torch.movedim(yh, -1, self.dims[0])
            return torch.stack([getattr(self.estimators_[i], method)(Xm[..., i], *args) for i in range(Xm.shape[-1])], self.dims[0])
        Xm, ym, eff = self._moved(X, y)
        fn = getattr(_Sliding_torch(estimator=self.estimator, dims=eff[1:], n_jobs=None, top=False), method) if len(self.dims) > 1 \
            else self.estimator
        Z = torch.stack([fn(Xm[..., i], ym[..., j], *args) for i, j in self._pairs(Xm, ym)], self.dims[0] - 1)
        return Z.swapaxes(0, 1) if self.top else Z

    def transform(self, X, y=None, *args):
        return self._apply('transform', X, y, *args)

    def fit_transform(self, X, y=None, *args):
        return self.fit(X, y, *args).transform(X, y, *args)

    def decision_function(self, X, y=None, *args):
        return self._apply('decision_function', X, y, *args)

    def predict(self, X, y=None, *args):
        return self._apply('predict', X, y, *args)

    def predict_proba(self, X, y=None, *args):
        return self._apply('predict_proba', X, y, *args)

    def score(self, X: torch.Tensor, y: torch.Tensor, metric: Optional[Union[metrics.Metric, Tuple[metrics.Metric]]] = None):
        if metric is None:
            metric = self.estimators_[0].metric_
        return metrics.score(self, metric, X, y)

    def collect(self, attr: str) -> torch.Tensor:
        if isinstance(self.estimators_[0], _Sliding_torch):
            return torch.stack([e.collect(attr) for e in self.estimators_])
        if hasattr(self.estimators_[0], attr) == False:
            raise AttributeError(f'Estimator of type `{type(self.estimators_[0])}` does not have attribute `{attr}`.')
        return torch.stack([getattr(e, attr) for e in self.estimators_])

    def clone(self):
        return _Sliding_torch(estimator=self.estimator, dims=self.dims, n_jobs=self.n_jobs, top=self.top, verbose=self.verbose)


class Sliding(sklearn.base.BaseEstimator):
    """Dispatcher (sliding.py:1026-1060).  numpy ``dims`` select the numpy backend in the reference;
    here every spelling of ``dims`` gives the torch/CUDA implementation."""

    def __new__(cls, estimator: Union[Callable, sklearn.base.BaseEstimator], dims: Union[int, tuple, list, np.ndarray, torch.Tensor] = -1,
                n_jobs: Union[int, None] = None, top: bool = True, verbose: bool = False):
        if isinstance(dims, int):
            dims = (dims,)
        elif isinstance(dims, list):
            dims = tuple(dims)
        elif isinstance(dims, np.ndarray):
            dims = tuple(int(d) for d in dims.flatten())
        elif isinstance(dims, torch.Tensor):
            dims = tuple(dims.flatten().tolist())
        elif not isinstance(dims, tuple):
            raise ValueError(f'`dims` must be an integer, tuple, list, numpy array or torch tensor, but got {type(dims)}.')
        if (isinstance(estimator, sklearn.base.BaseEstimator) == False) & (callable(estimator) == False):
            raise ValueError(f'`estimator` must be sklearn.base.BaseEstimator or Callable, but got {type(estimator)}.')
        return _Sliding_torch(estimator, dims=dims, n_jobs=n_jobs, top=top, verbose=verbose)
