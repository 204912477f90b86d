"""Standard scaler -- drop-in for ``mvpy.preprocessing.Scaler`` (mvpy/preprocessing/scaler.py:232-672).

Statistics and transforms run in the ``mvb_scaler_fit`` / ``mvb_scaler_apply`` CUDA kernels; this
file only maps the reference's ``dims`` semantics onto the kernels' [R][K] view."""
from __future__ import annotations

from typing import Any, Optional, Tuple, Union

import numpy as np
import sklearn.base
import torch

from .. import _lib


def _normalise_dims(dims, ndim: int) -> Tuple[int, ...]:
    if dims is None:
        dims = (0,)
    if isinstance(dims, torch.Tensor):
        dims = dims.flatten().tolist()
    if isinstance(dims, (int, np.integer)):
        dims = (int(dims),)
    return tuple(int(d) for d in dims)


class _Scaler_torch(sklearn.base.BaseEstimator):
    """z-scoring over ``dims`` (default: dim 0 = trials, i.e. one mean/scale per (feature, time));
    unbiased variance; zero or NaN scale becomes 1; transforms never touch their input
    (mvpy/preprocessing/scaler.py:301-428).  Attributes: ``shape_, mean_, var_, scale_`` (keepdim shapes)."""

    def __init__(self, with_mean: bool = True, with_std: bool = True, dims: Union[list, tuple, int, None] = None):
        self.dims = self.dims_ = dims
        self.with_mean = self.with_mean_ = with_mean
        self.with_std = self.with_std_ = with_std
        self.shape_ = None
        self.mean_ = None
        self.var_ = None
        self.scale_ = None

    # -- layout helpers -------------------------------------------------------------------------
    def _split(self, ndim: int):
        red = tuple(sorted(d % ndim for d in self.dims_))
        keep = tuple(i for i in range(ndim) if i not in red)
        return red, keep

    def fit(self, X: torch.Tensor, *args: Any, sample_weight: Optional[torch.Tensor] = None):
        if not isinstance(X, torch.Tensor):
            raise TypeError(f'`X` must be a torch tensor, but got {type(X)}.')
        self.dims_ = _normalise_dims(self.dims_, X.ndim)
        if sample_weight is not None:
            if not isinstance(sample_weight, torch.Tensor):
                raise TypeError(f'`sample_weight` must be a torch tensor, but got {type(sample_weight)}.')
            if sample_weight.shape[0] != X.shape[0]:
                raise ValueError(f'`sample_weight` must match `X` in first dimension, but got {sample_weight.shape} and {X.shape}.')
        self.shape_ = X.shape
        self.mean_ = self.var_ = self.scale_ = None
        out_dev = X.device
        Xd = _lib.to_device(X)
        if sample_weight is not None:
            # weighted statistics (scaler.py:349-353): rarely used, kept as device-side tensor algebra
            w = _lib.to_device(sample_weight).reshape((X.shape[0],) + (1,) * (X.ndim - 1)).to(Xd.dtype)
            n = (w != 0.0).sum(dim=0, keepdim=True)
            sw = w.sum(dim=self.dims_, keepdim=True)
            mean = (Xd * w).sum(dim=self.dims_, keepdim=True) / sw
            var = ((Xd - mean) ** 2 * w).sum(dim=self.dims_, keepdim=True) / ((n - 1) / n * sw)
            scale = torch.sqrt(var)
            scale[(scale == 0.0) | torch.isnan(scale)] = 1.0
        else:
            red, keep = self._split(X.ndim)
            Xp = Xd.permute(*red, *keep).contiguous()
            R = int(np.prod([X.shape[d] for d in red])) if red else 1
            K = int(np.prod([X.shape[d] for d in keep])) if keep else 1
            mean, var, scale = _lib.scaler_fit(Xp.reshape(R, K))
            kshape = tuple(1 if i in red else X.shape[i] for i in range(X.ndim))
            mean, var, scale = mean.reshape(kshape), var.reshape(kshape), scale.reshape(kshape)
        self.mean_, self.var_, self.scale_ = mean.to(out_dev), var.to(out_dev), scale.to(out_dev)
        return self

    def _apply(self, X: torch.Tensor, inverse: bool) -> torch.Tensor:
        if (self.with_mean_ and self.mean_ is None) or (self.with_std_ and self.scale_ is None):
            raise ValueError('The scaler has not been fitted yet.')
        if not isinstance(X, torch.Tensor):
            raise TypeError(f'`X` must be a torch tensor, but got {type(X)}.')
        out_dev = X.device
        Xd = _lib.to_device(X)
        red, keep = self._split(X.ndim)
        mean, scale = _lib.to_device(self.mean_).to(Xd.dtype), _lib.to_device(self.scale_).to(Xd.dtype)
        lead_ok = red == tuple(range(len(red)))       # reduced dims lead -> X is already [R][K]
        stats_match = all(self.mean_.shape[i] == X.shape[i] for i in keep)
        if not stats_match:
            raise RuntimeError(f'The size of `X` {tuple(X.shape)} does not match the fitted statistics {tuple(self.mean_.shape)}.')
        Xc = Xd.contiguous() if lead_ok else Xd.permute(*red, *keep).contiguous()
        K = int(np.prod([X.shape[d] for d in keep])) if keep else 1
        R = Xc.numel() // K
        out = _lib.scaler_apply(Xc.reshape(R, K), mean.reshape(-1).contiguous(), scale.reshape(-1).contiguous(),
                                self.with_mean_, self.with_std_, inverse)
        out = out.reshape(Xc.shape)
        if not lead_ok:
            inv = np.argsort(red + keep)
            out = out.permute(*inv).contiguous()
        return out.to(out_dev)

    def transform(self, X: torch.Tensor, *args: Any) -> torch.Tensor:
        return self._apply(X, inverse=False)

    def inverse_transform(self, X: torch.Tensor, *args: Any) -> torch.Tensor:
        return self._apply(X, inverse=True)

    def fit_transform(self, X: torch.Tensor, *args: Any, sample_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.fit(X, *args, sample_weight=sample_weight).transform(X, *args)

    def clone(self):
        return _Scaler_torch(with_mean=self.with_mean, with_std=self.with_std_, dims=self.dims_)


class Scaler(sklearn.base.BaseEstimator):
    """Dispatcher with the reference's surface (scaler.py:461-672): ``fit`` returns the fitted backend
    object; ``to_torch()`` gives the pipeline-ready estimator.  numpy inputs are rejected -- this build
    serves the torch / CUDA side only."""

    def __init__(self, with_mean: bool = True, with_std: bool = True, dims: Union[list, tuple, int, None] = None):
        self.with_mean = with_mean
        self.with_std = with_std
        self.dims = dims

    def _get_estimator(self, X, *args):
        if isinstance(X, torch.Tensor):
            return _Scaler_torch
        raise TypeError(f'`X` must be a torch.Tensor for the B200 build, but got {type(X)}.')

    def _make(self, X):
        return self._get_estimator(X)(with_mean=self.with_mean, with_std=self.with_std, dims=self.dims)

    def fit(self, X, *args, sample_weight=None):
        return self._make(X).fit(X, *args, sample_weight=sample_weight)

    def transform(self, X, *args):
        return self._make(X).transform(X, *args)

    def inverse_transform(self, X, *args):
        return self._make(X).inverse_transform(X, *args)

    def fit_transform(self, X, *args, sample_weight=None):
        return self._make(X).fit_transform(X, *args, sample_weight=sample_weight)

    def to_torch(self):
        return _Scaler_torch(with_mean=self.with_mean, with_std=self.with_std, dims=self.dims)

    def to_numpy(self):
        raise NotImplementedError('The B200 build has no numpy backend (no CPU fallback).')

    def clone(self):
        return Scaler(with_mean=self.with_mean, with_std=self.with_std, dims=self.dims)

    def copy(self):
        return self.clone()
