"""ctypes binding of libmvb.so (include/mvb.h) and thin tensor-level wrappers.

This is the only place that touches the C-ABI.  Everything here enqueues CUDA work on torch's
current stream; torch is used for device memory (outputs, workspaces) and streams only.
There is no CPU implementation: importing works without a GPU (so host-side logic can be tested),
but any compute call raises if the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes
import threading
import os
from ctypes import POINTER, c_char_p, c_int, c_int32, c_int64, c_size_t, c_void_p
from typing import Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmvb.so')

F32, F64 = 0, 1


class MvbError(RuntimeError):
    pass


class mvb_design(ctypes.Structure):
    _fields_ = [('xp', c_void_p), ('trial_stride', c_int64), ('feat_stride', c_int64), ('n_time', c_int),
                ('n_lag', c_int), ('trial_idx', c_void_p), ('n_trials', c_int), ('feat_idx', c_void_p),
                ('n_feat', c_int), ('dtype', c_int)]


class mvb_strided(ctypes.Structure):
    _fields_ = [('base', c_void_p), ('sample_stride', c_int64), ('col_stride', c_int64), ('batch_stride', c_int64)]


class mvb_operand(ctypes.Structure):
    _fields_ = [('base', c_void_p), ('roff', c_void_p), ('koff', c_void_p), ('rows', c_int), ('contig', c_int)]


# every symbol declared in include/mvb.h (tests/test_abi.py checks the list against the header)
_SIGNATURES = {
    'mvb_version': (c_int, []),
    'mvb_last_error': (c_char_p, []),
    'mvb_launch_count': (c_int64, []),
    'mvb_profile_reset': (None, []),
    'mvb_profile_disable': (None, []),
    'mvb_profile_set': (c_int, [c_int]),
    'mvb_launch_count_add': (None, [c_int64]),
    'mvb_profile_read': (c_int, [POINTER(ctypes.c_double), POINTER(ctypes.c_double), POINTER(ctypes.c_double),
                                 POINTER(c_int64), POINTER(c_int64), c_int]),
    'mvb_profile_name': (c_char_p, [c_int]),
    'mvb_scaler_fit': (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'mvb_scaler_apply': (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'mvb_trf_pad': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'mvb_transpose': (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    'mvb_trf_shift': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'mvb_contract_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'mvb_contract': (c_int, [POINTER(mvb_operand), POINTER(mvb_operand), c_int, c_int, c_int, c_int, c_void_p, c_int64,
                             c_int, c_void_p, c_size_t, c_void_p]),
    'mvb_trf_stats_workspace_bytes': (c_size_t, [POINTER(mvb_design), c_int]),
    'mvb_trf_stats': (c_int, [POINTER(mvb_design), c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_size_t, c_void_p]),
    'mvb_gather_stats': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    'mvb_eigh_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'mvb_eigh': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    'mvb_ridge_solve_workspace_bytes': (c_size_t, [POINTER(mvb_design), c_int, c_int]),
    'mvb_ridge_solve': (c_int, [POINTER(mvb_design), c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_size_t, c_void_p]),
    'mvb_band_padded': (c_int, [c_int]),
    'mvb_band_reduce_workspace_bytes': (c_size_t, [c_int]),
    'mvb_band_reduce': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'mvb_band_reduce_counters': (c_int, [c_void_p, c_int, c_void_p]),
    'mvb_trf_predict_workspace_bytes': (c_size_t, [POINTER(mvb_design), c_int]),
    'mvb_trf_predict': (c_int, [POINTER(mvb_design), c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'mvb_rank': (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    'mvb_accuracy': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    'mvb_roc_auc': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    'mvb_pairwise_coupling_workspace_bytes': (c_size_t, [c_int64, c_int, c_int]),
    'mvb_pairwise_coupling': (c_int, [c_void_p, c_int64, c_int, c_int, ctypes.c_double, ctypes.c_double, c_int, c_void_p, c_void_p, c_size_t,
                                      c_void_p]),
    'mvb_rf_assemble': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'mvb_sum_slabs': (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'mvb_penalised_solve': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'mvb_score': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    'mvb_dense_ridge_batched_max_cols': (c_int, []),
    'mvb_dense_ridge_batched_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    'mvb_dense_ridge_batched': (c_int, [POINTER(mvb_strided), POINTER(mvb_strided), c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                        c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_size_t, c_void_p]),
    'mvb_dense_predict_batched': (c_int, [POINTER(mvb_strided), c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int64,
                                          c_int64, c_int64, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libmvb.so; loud failure if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MvbError(f'{LIB_PATH} not found: build it with `python -m mvpy_b200.build` '
                       '(nvcc, sm_100a).  mvpy_b200 has no CPU or PyTorch fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise MvbError(f'{what} failed (rc={rc}): {load().mvb_last_error().decode()}')


def launch_count() -> int:
    return int(load().mvb_launch_count())


def profile_reset() -> None:
    load().mvb_profile_reset()


def profile_read(disable: bool = True) -> dict:
    """Per kernel class: ms / flops / n are ESTIMATES for all launches of the class (timed subset scaled by flops),
    plus the raw timed figures.  Synchronises the recorded events."""
    lib = load()
    nc = 16
    ms, ft, fa = (ctypes.c_double * nc)(), (ctypes.c_double * nc)(), (ctypes.c_double * nc)()
    nt, na = (c_int64 * nc)(), (c_int64 * nc)()
    n = lib.mvb_profile_read(ms, ft, fa, nt, na, nc)
    out = {}
    for c in range(n):
        if na[c] == 0:
            continue
        scale = fa[c] / ft[c] if ft[c] > 0 else 0.0
        out[lib.mvb_profile_name(c).decode()] = dict(ms=ms[c] * scale, flops=fa[c], n=int(na[c]), ms_timed=ms[c],
                                                     flops_timed=ft[c], n_timed=int(nt[c]))
    if disable:
        lib.mvb_profile_disable()
    return out


# ------------------------------------------------------------------------------ shared-design context
class _FoldContext(threading.local):
    """What the predictor-set drivers (Hierarchical / Shapley) and the Validator know and the estimator at the end of
    the chain does not: that the data it is about to be fitted on is a feature subset (``idx`` along ``dim``) of a larger
    array ``parent`` restricted to the training trials ``train``, after the preprocessing steps of ``model``.  With that,
    ``TimeDelayed.fit`` computes the fold's full Gram / cross-moment once and slices it for every predictor set
    (``shared`` is the per-driver cache)."""
    parent = None       # dict(X=..., dim=..., idx=..., shared={...}) set by fit_validator_
    train = None        # training indices of the current fold, set by fit_model_
    model = None        # the model clone being fitted in the current fold, set by fit_model_
    train_key = None    # host-side identity of ``train`` (validator.fold_key), set by fit_model_ when the driver computed one


fold_ctx = _FoldContext()


# ------------------------------------------------------------------------------------------- helpers
def compute_device() -> torch.device:
    """CUDA device used for host (CPU) inputs."""
    if not torch.cuda.is_available():
        raise MvbError('mvpy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.')
    return torch.device('cuda', torch.cuda.current_device())


def to_device(x: torch.Tensor) -> torch.Tensor:
    """Move a host tensor to the compute device (non-blocking from pinned memory); CUDA tensors pass."""
    if x.is_cuda:
        return x
    return x.to(compute_device(), non_blocking=True)


def stage(x: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """Driver-level staging (Validator / Hierarchical / Shapley): host tensors go to the GPU once when one
    is present.  Without a GPU the tensor is returned as is -- the drivers are pure host logic and stay
    usable with user-supplied models; the CUDA estimators themselves still raise in ``to_device``."""
    if x is None or x.is_cuda or not torch.cuda.is_available():
        return x
    return x.to(compute_device(), non_blocking=True)


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float64:
        return F64
    raise TypeError(f'mvpy_b200 kernels support float32 / float64 tensors, got {t.dtype}')


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else c_void_p(t.data_ptr())


def _stream(t: torch.Tensor):
    return c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


class Design:
    """Python mirror of ``mvb_design``: the implicit lag-expanded design over a padded stimulus array.

    xp: (n_all_trials, C, Tp) contiguous CUDA tensor (edge-replicated by ``trf_pad``) or, for dense 2-D
    designs, a (1, p, n) "series" array with n_lag = 1.
    """

    def __init__(self, xp: torch.Tensor, n_time: int, n_lag: int, trial_idx: Optional[torch.Tensor] = None,
                 feat_idx: Optional[torch.Tensor] = None):
        assert xp.is_cuda and xp.is_contiguous() and xp.ndim == 3
        self.xp = xp
        self.n_time, self.n_lag = int(n_time), int(n_lag)
        self.trial_idx = None if trial_idx is None else trial_idx.to(device=xp.device, dtype=torch.int32).contiguous()
        self.feat_idx = None if feat_idx is None else feat_idx.to(device=xp.device, dtype=torch.int32).contiguous()
        self.n_trials = xp.shape[0] if self.trial_idx is None else int(self.trial_idx.numel())
        self.n_feat = xp.shape[1] if self.feat_idx is None else int(self.feat_idx.numel())

    @property
    def n_rows(self) -> int:
        return self.n_trials * self.n_time

    @property
    def n_cols(self) -> int:
        return self.n_feat * self.n_lag

    def subset(self, trial_idx=None, feat_idx=None) -> 'Design':
        return Design(self.xp, self.n_time, self.n_lag, trial_idx if trial_idx is not None else self.trial_idx,
                      feat_idx if feat_idx is not None else self.feat_idx)

    def c_struct(self) -> mvb_design:
        xp = self.xp
        return mvb_design(xp.data_ptr(), xp.stride(0), xp.stride(1), self.n_time, self.n_lag,
                          None if self.trial_idx is None else self.trial_idx.data_ptr(), self.n_trials,
                          None if self.feat_idx is None else self.feat_idx.data_ptr(), self.n_feat, dtype_code(xp))


# --------------------------------------------------------------------------------------------- ops
def scaler_fit(X2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """X2: (R, K) contiguous -> mean, var, scale (K,)."""
    R, K = X2.shape
    mean, var, scale = (torch.empty(K, dtype=X2.dtype, device=X2.device) for _ in range(3))
    check(load().mvb_scaler_fit(_ptr(X2), R, K, dtype_code(X2), _ptr(mean), _ptr(var), _ptr(scale), _stream(X2)), 'mvb_scaler_fit')
    return mean, var, scale


def scaler_apply(X2: torch.Tensor, mean, scale, with_mean: bool, with_std: bool, inverse: bool = False) -> torch.Tensor:
    R, K = X2.shape
    out = torch.empty_like(X2)
    check(load().mvb_scaler_apply(_ptr(X2), R, K, dtype_code(X2), _ptr(mean), _ptr(scale), int(with_mean), int(with_std),
                                  int(inverse), _ptr(out), _stream(X2)), 'mvb_scaler_apply')
    return out


def trf_pad(X: torch.Tensor, left: int, right: int) -> torch.Tensor:
    N, C, T = X.shape
    out = torch.empty((N, C, T + left + right), dtype=X.dtype, device=X.device)
    check(load().mvb_trf_pad(_ptr(X), N, C, T, left, right, dtype_code(X), _ptr(out), _stream(X)), 'mvb_trf_pad')
    return out


def shift_design(design: 'Design'):
    """Centring before multiplying: subtract from every feature of the padded stimulus its mean over the design's trials.
    Returns (design on the shifted array, shift (C_all,)); column (c, w) of the shifted design is the original minus
    shift[c], which leaves the centred Gram / cross-moment unchanged (see ``mvb_trf_shift``)."""
    xp = design.xp
    N, C, Tp = xp.shape
    shift = torch.empty(C, dtype=xp.dtype, device=xp.device)
    out = torch.empty_like(xp)
    check(load().mvb_trf_shift(_ptr(xp), N, C, Tp, _ptr(design.trial_idx), design.n_trials, dtype_code(xp), _ptr(shift), _ptr(out),
                               _stream(xp)), 'mvb_trf_shift')
    return Design(out, design.n_time, design.n_lag, design.trial_idx, design.feat_idx), shift


def shift_columns(design: 'Design', shift: torch.Tensor) -> torch.Tensor:
    """The per-column (c, w) view of a per-feature shift: (p,)."""
    s = shift if design.feat_idx is None else shift[design.feat_idx.long()]
    return s.repeat_interleave(design.n_lag)


def transpose(X2: torch.Tensor) -> torch.Tensor:
    R, C = X2.shape
    out = torch.empty((C, R), dtype=X2.dtype, device=X2.device)
    check(load().mvb_transpose(_ptr(X2), R, C, dtype_code(X2), _ptr(out), _stream(X2)), 'mvb_transpose')
    return out


def trf_stats(design: Design, y: torch.Tensor, fit_intercept: bool = True):
    """Centred Gram / cross-moment / means of the training rows of ``design`` against y (N, F, T)."""
    F = y.shape[1]
    p = design.n_cols
    dev, dt = design.xp.device, design.xp.dtype
    G = torch.empty((p, p), dtype=dt, device=dev)
    XtY = torch.empty((p, F), dtype=dt, device=dev)
    mu = torch.empty(p, dtype=dt, device=dev)
    ybar = torch.empty(F, dtype=dt, device=dev)
    cs = design.c_struct()
    lib = load()
    ws = _workspace(lib.mvb_trf_stats_workspace_bytes(ctypes.byref(cs), F), dev)
    check(lib.mvb_trf_stats(ctypes.byref(cs), _ptr(y), F, int(fit_intercept), _ptr(G), _ptr(XtY), _ptr(mu), _ptr(ybar),
                            _ptr(ws), ws.numel(), _stream(y)), 'mvb_trf_stats')
    return G, XtY, mu, ybar


def gather_stats(G: torch.Tensor, XtY: torch.Tensor, mu: torch.Tensor, feat_idx: torch.Tensor, n_lag: int):
    """(G, XtY, mu) of the predictor subset ``feat_idx`` sliced out of the full-predictor statistics (one kernel)."""
    fi = feat_idx.to(device=G.device, dtype=torch.int32).contiguous()
    ps, F = int(fi.numel()) * n_lag, XtY.shape[1]
    Gs = torch.empty((ps, ps), dtype=G.dtype, device=G.device)
    Bs = torch.empty((ps, F), dtype=G.dtype, device=G.device)
    ms = torch.empty(ps, dtype=G.dtype, device=G.device)
    check(load().mvb_gather_stats(_ptr(G), _ptr(XtY), _ptr(mu), G.shape[0], F, _ptr(fi), int(fi.numel()), n_lag, dtype_code(G), _ptr(Gs),
                                  _ptr(Bs), _ptr(ms), _stream(G)), 'mvb_gather_stats')
    return Gs, Bs, ms


def ridge_solve(design: Design, y: torch.Tensor, G, XtY, mu, ybar, alphas: torch.Tensor, fit_intercept: bool,
                alpha_per_target: bool, want_loss: bool = False):
    """LOO sweep from sufficient statistics (G is overwritten).  Returns coef (F,p), intercept (F,),
    alpha (1 or F), best (int32), loss (A,F) or None."""
    F = y.shape[1]
    p, A = design.n_cols, int(alphas.numel())
    dev, dt = design.xp.device, design.xp.dtype
    coef = torch.empty((F, p), dtype=dt, device=dev)
    icpt = torch.empty(F, dtype=dt, device=dev)
    nb = F if alpha_per_target else 1
    alpha_out = torch.empty(nb, dtype=dt, device=dev)
    best = torch.empty(nb, dtype=torch.int32, device=dev)
    loss = torch.empty((A, F), dtype=dt, device=dev) if want_loss else None
    cs = design.c_struct()
    lib = load()
    ws = _workspace(lib.mvb_ridge_solve_workspace_bytes(ctypes.byref(cs), F, A), dev)
    check(lib.mvb_ridge_solve(ctypes.byref(cs), _ptr(y), F, _ptr(G), _ptr(XtY), _ptr(mu), _ptr(ybar), _ptr(alphas), A,
                              int(fit_intercept), int(alpha_per_target), _ptr(coef), _ptr(icpt), _ptr(alpha_out), _ptr(best),
                              _ptr(loss), _ptr(ws), ws.numel(), _stream(y)), 'mvb_ridge_solve')
    return coef, icpt, alpha_out, best, loss


def trf_predict(design: Design, coef: torch.Tensor, intercept: Optional[torch.Tensor]) -> torch.Tensor:
    """(n_trials, F, T) predictions for the (selected) trials of ``design``."""
    F = coef.shape[0]
    dev, dt = design.xp.device, design.xp.dtype
    out = torch.empty((design.n_trials, F, design.n_time), dtype=dt, device=dev)
    cs = design.c_struct()
    lib = load()
    ws = _workspace(lib.mvb_trf_predict_workspace_bytes(ctypes.byref(cs), F), dev)
    check(lib.mvb_trf_predict(ctypes.byref(cs), _ptr(coef), _ptr(intercept), F, _ptr(out), _ptr(ws), ws.numel(), _stream(coef)),
          'mvb_trf_predict')
    return out


def score(y2: torch.Tensor, yh2: torch.Tensor, yb: Optional[torch.Tensor], want_r2: bool, want_pearson: bool):
    """y2, yh2: (R, K) contiguous; yb: (K,) or None -> (r2 (K,) | None, pearson (K,) | None)."""
    R, K = y2.shape
    r2 = torch.empty(K, dtype=y2.dtype, device=y2.device) if want_r2 else None
    pr = torch.empty(K, dtype=y2.dtype, device=y2.device) if want_pearson else None
    check(load().mvb_score(_ptr(y2), _ptr(yh2), _ptr(yb), R, K, dtype_code(y2), _ptr(r2), _ptr(pr), _stream(y2)), 'mvb_score')
    return r2, pr


def accuracy_rows(y2: torch.Tensor, yh2: torch.Tensor) -> torch.Tensor:
    """y2, yh2: (R, K) contiguous, same dtype -> fraction of equal entries along R, (K,)."""
    R, K = y2.shape
    out = torch.empty(K, dtype=y2.dtype, device=y2.device)
    check(load().mvb_accuracy(_ptr(y2), _ptr(yh2), R, K, dtype_code(y2), _ptr(out), _stream(y2)), 'mvb_accuracy')
    return out


def roc_auc_rows(pos2: torch.Tensor, score2: torch.Tensor) -> torch.Tensor:
    """pos2 (0/1), score2: (R, K) contiguous, same dtype -> area under the ROC curve along R, (K,)."""
    R, K = pos2.shape
    out = torch.empty(K, dtype=pos2.dtype, device=pos2.device)
    check(load().mvb_roc_auc(_ptr(pos2), _ptr(score2), R, K, dtype_code(pos2), _ptr(out), _stream(pos2)), 'mvb_roc_auc')
    return out


def pairwise_coupling(p_pair: torch.Tensor, max_iter: int, eps: float, tol: float) -> torch.Tensor:
    """p_pair (N, K, K) pairwise probabilities -> (N, K) coupled class probabilities (Wu-Lin, reference semantics)."""
    p_pair = p_pair.contiguous()
    N, K, _ = p_pair.shape
    out = torch.empty((N, K), dtype=p_pair.dtype, device=p_pair.device)
    lib = load()
    ws = _workspace(lib.mvb_pairwise_coupling_workspace_bytes(N, K, max_iter), p_pair.device)
    check(lib.mvb_pairwise_coupling(_ptr(p_pair), N, K, int(max_iter), float(eps), float(tol), dtype_code(p_pair), _ptr(out), _ptr(ws),
                                    ws.numel(), _stream(p_pair)), 'mvb_pairwise_coupling')
    return out


def rank_rows(x2: torch.Tensor) -> torch.Tensor:
    """x2: (R, K) contiguous -> average ranks along R for every column, same shape / dtype."""
    R, K = x2.shape
    out = torch.empty_like(x2)
    check(load().mvb_rank(_ptr(x2), R, K, dtype_code(x2), _ptr(out), _stream(x2)), 'mvb_rank')
    return out


def rf_assemble(G: torch.Tensor, Xc: torch.Tensor, s_min: int, s_max: int, edge_correction: bool) -> torch.Tensor:
    """G (N, p, p) in kernel lag order + centred epochs Xc (N, C, T) -> cov_ (N, p, p) in the reference's layout."""
    N, C, Tn = Xc.shape
    cov = torch.empty_like(G)
    check(load().mvb_rf_assemble(_ptr(G), _ptr(Xc), N, C, s_max - s_min, Tn, s_min, s_max, int(edge_correction), dtype_code(G),
                                 _ptr(cov), _stream(G)), 'mvb_rf_assemble')
    return cov


def sum_slabs(A: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """A (N, ...) contiguous -> sum over the slabs listed in idx."""
    idx = idx.to(device=A.device, dtype=torch.int32).contiguous()
    out = torch.empty(A.shape[1:], dtype=A.dtype, device=A.device)
    check(load().mvb_sum_slabs(_ptr(A), out.numel(), _ptr(idx), idx.numel(), dtype_code(A), _ptr(out), _stream(A)), 'mvb_sum_slabs')
    return out


def penalised_solve(G: torch.Tensor, reg: torch.Tensor, alphas: torch.Tensor, rhs: torch.Tensor):
    """(G + alpha reg) x = rhs for every alpha: solutions (S, p, F) and info (S,) int32 (0 = ok)."""
    p, F, S = G.shape[0], rhs.shape[1], int(alphas.numel())
    mats = torch.empty((S, p, p), dtype=G.dtype, device=G.device)
    sol = torch.empty((S, p, F), dtype=G.dtype, device=G.device)
    info = torch.zeros(S, dtype=torch.int32, device=G.device)
    check(load().mvb_penalised_solve(_ptr(G), _ptr(reg), _ptr(alphas), S, _ptr(rhs), p, F, dtype_code(G), _ptr(mats), _ptr(sol),
                                     _ptr(info), _stream(G)), 'mvb_penalised_solve')
    return sol, info


def pinv_symmetric(S: torch.Tensor) -> torch.Tensor:
    """Pseudo-inverse of a small symmetric positive semi-definite matrix (the (f x f) target covariance of the Haufe
    patterns) from the library's Jacobi eigensolver; eigenvalues below f * eps * max are dropped like
    ``torch.linalg.pinv`` does with singular values.  Stream-ordered (no host synchronisation), so it can be captured."""
    f = S.shape[0]
    Vt, lam = eigh(S.clone()[None].contiguous())
    Vt, lam = Vt[0], lam[0]
    keep = lam > lam.max() * (f * torch.finfo(S.dtype).eps)
    inv = torch.where(keep, 1.0 / torch.where(keep, lam, torch.ones_like(lam)), torch.zeros_like(lam))
    return (Vt * inv[:, None]).T @ Vt


def eigh(A: torch.Tensor):
    """Batched symmetric eigensolve.  A: (b, p, p) (destroyed) -> Vt (b, p, p) rows = eigenvectors, lam (b, p)."""
    b, p, _ = A.shape
    Vt = torch.empty_like(A)
    lam = torch.empty((b, p), dtype=A.dtype, device=A.device)
    lib = load()
    ws = _workspace(lib.mvb_eigh_workspace_bytes(b, p, dtype_code(A)), A.device)
    check(lib.mvb_eigh(_ptr(A), _ptr(Vt), _ptr(lam), b, p, dtype_code(A), _ptr(ws), ws.numel(), _stream(A)), 'mvb_eigh')
    return Vt, lam


def band_reduce(G: torch.Tensor, counters: bool = False):
    """G (p, p) fp32 symmetric -> Qt (P, P), Tdiag, Tsub (P/128, 128, 128) with pad(G) = Qt^T T Qt."""
    p = G.shape[0]
    lib = load()
    P = lib.mvb_band_padded(p)
    Qt = torch.zeros((P, P), dtype=G.dtype, device=G.device)
    Td = torch.zeros((P // 128, 128, 128), dtype=G.dtype, device=G.device)
    Ts = torch.zeros_like(Td)
    ws = _workspace(lib.mvb_band_reduce_workspace_bytes(p), G.device)
    check(lib.mvb_band_reduce(_ptr(G), p, _ptr(Qt), _ptr(Td), _ptr(Ts), _ptr(ws), ws.numel(), _stream(G)), 'mvb_band_reduce')
    if counters:
        out = (ctypes.c_uint * 2)()
        check(lib.mvb_band_reduce_counters(_ptr(ws), p, out), 'mvb_band_reduce_counters')
        return Qt, Td, Ts, (int(out[0]), int(out[1]))
    return Qt, Td, Ts


def matmul_nt(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """A (M, K) @ B (N, K)^T on the library's contraction kernel (both operands row-major, k contiguous)."""
    A, B = A.contiguous(), B.contiguous()
    M, K = A.shape
    N = B.shape[0]
    dev = A.device
    a_r = torch.arange(M, device=dev, dtype=torch.int64) * K
    b_r = torch.arange(N, device=dev, dtype=torch.int64) * K
    k_o = torch.arange(K, device=dev, dtype=torch.int64)
    vec = 2 if (A.dtype == torch.float32 and K % 4 == 0 and A.data_ptr() % 16 == 0 and B.data_ptr() % 16 == 0) else 1
    return contract(A, a_r, k_o, vec, B, b_r, k_o, vec, M, N, K)


def covariance(Y: torch.Tensor) -> torch.Tensor:
    """Unbiased covariance of the columns of Y (n, f) -> (f, f), like ``torch.cov(Y.T)``: centred columns, then one
    contraction over the samples on the library kernel."""
    n = Y.shape[0]
    Yt = transpose((Y - Y.mean(0, keepdim=True)).contiguous())          # (f, n), samples contiguous
    return matmul_nt(Yt, Yt) / float(n - 1)


def contract(a_base, a_roff, a_koff, a_contig, b_base, b_roff, b_koff, b_contig, M, N, K, symmetric=False) -> torch.Tensor:
    """Generic separable-gather contraction C[M,N] = sum_k A(m,k) B(n,k) (tests, patterns)."""
    C = torch.empty((M, N), dtype=a_base.dtype, device=a_base.device)
    A = mvb_operand(a_base.data_ptr(), a_roff.data_ptr(), a_koff.data_ptr(), M, a_contig)
    B = mvb_operand(b_base.data_ptr(), b_roff.data_ptr(), b_koff.data_ptr(), N, b_contig)
    lib = load()
    ws = _workspace(lib.mvb_contract_workspace_bytes(M, N, K, dtype_code(a_base)), a_base.device)
    check(lib.mvb_contract(ctypes.byref(A), ctypes.byref(B), M, N, K, dtype_code(a_base), _ptr(C), N, int(symmetric), _ptr(ws),
                           ws.numel(), _stream(a_base)), 'mvb_contract')
    return C


DENSE_BATCH_WS_LIMIT = 12 << 30      # bytes of scratch per mvb_dense_ridge_batched call; larger stacks are fitted in chunks


def _strided(t3: torch.Tensor) -> mvb_strided:
    return mvb_strided(t3.data_ptr(), t3.stride(0), t3.stride(1), t3.stride(2))


def dense_ridge_batched_ok(n: int, p: int, F: int, A: int, dtype) -> bool:
    """Shapes the batched small-problem route takes (float32, p <= 320, more samples than columns)."""
    if dtype != torch.float32:
        return False
    return p <= load().mvb_dense_ridge_batched_max_cols() and p + 1 < n <= 65535 and A * F <= 4096


def dense_ridge_batched(X3: torch.Tensor, Y3: torch.Tensor, sample_idx: Optional[torch.Tensor], alphas: torch.Tensor, fit_intercept: bool,
                        alpha_per_target: bool, want_loss: bool = True, want_gram: bool = False):
    """B independent LOO ridge fits in one call (``mvb_dense_ridge_batched``).  X3: (N, p, B) and Y3: (N, F, B) float32 CUDA
    tensors of ANY strides (problem b is X3[sample_idx, :, b] -> Y3[sample_idx, :, b]; nothing is copied here).
    Returns dict(coef (B, F, p), intercept (B, F), alpha (B, F) or (B,), best, loss (B, A, F) | None, gram (B, p, p) | None,
    status (B,) int32: non-zero = refit that problem on the general route)."""
    N, p, B = X3.shape
    F = Y3.shape[1]
    A = int(alphas.numel())
    dev = X3.device
    si = None if sample_idx is None else sample_idx.to(device=dev, dtype=torch.int32).contiguous()
    n = N if si is None else int(si.numel())
    lib = load()
    nb = B if alpha_per_target else 1
    coef = torch.empty((B, F, p), dtype=torch.float32, device=dev)
    icpt = torch.empty((B, F), dtype=torch.float32, device=dev)
    alpha = torch.empty((B, F) if alpha_per_target else (B,), dtype=torch.float32, device=dev)
    best = torch.empty((B, F) if alpha_per_target else (B,), dtype=torch.int32, device=dev)
    loss = torch.empty((B, A, F), dtype=torch.float32, device=dev) if want_loss else None
    gram = torch.empty((B, p, p), dtype=torch.float32, device=dev) if want_gram else None
    status = torch.empty(B, dtype=torch.int32, device=dev)
    al = alphas.to(device=dev, dtype=torch.float32).contiguous()
    per = lib.mvb_dense_ridge_batched_workspace_bytes(1, n, p, F, A)
    if per == 0:
        raise MvbError(f'mvb_dense_ridge_batched does not take n = {n}, p = {p}, F = {F}, A = {A}')
    chunk = max(1, min(B, 65535, int(DENSE_BATCH_WS_LIMIT // per)))
    ws = _workspace(lib.mvb_dense_ridge_batched_workspace_bytes(chunk, n, p, F, A), dev)
    st = _stream(X3)
    for b0 in range(0, B, chunk):
        bc = min(chunk, B - b0)
        xs, ys = _strided(X3[:, :, b0:b0 + bc]), _strided(Y3[:, :, b0:b0 + bc])
        check(lib.mvb_dense_ridge_batched(ctypes.byref(xs), ctypes.byref(ys), _ptr(si), n, p, F, bc, _ptr(al), A, int(fit_intercept),
                                          int(alpha_per_target), F32, _ptr(coef[b0:]), _ptr(icpt[b0:]), _ptr(alpha[b0:]), _ptr(best[b0:]),
                                          _ptr(loss[b0:]) if want_loss else None, _ptr(gram[b0:]) if want_gram else None,
                                          _ptr(status[b0:]), _ptr(ws), ws.numel(), st), 'mvb_dense_ridge_batched')
    return dict(coef=coef, intercept=icpt, alpha=alpha, best=best, loss=loss, gram=gram, status=status)


def dense_predict_batched(X3: torch.Tensor, coef: torch.Tensor, intercept: Optional[torch.Tensor]) -> torch.Tensor:
    """X3 (N, p, B) any strides, coef (B, F, p), intercept (B, F) | None -> predictions (N, F, B) (``mvb_dense_predict_batched``)."""
    N, p, B = X3.shape
    F = coef.shape[1]
    out = torch.empty((N, F, B), dtype=torch.float32, device=X3.device)
    xs = _strided(X3)
    coef = coef.contiguous()
    icpt = None if intercept is None else intercept.contiguous()
    check(load().mvb_dense_predict_batched(ctypes.byref(xs), N, p, B, _ptr(coef), _ptr(icpt), F, F32, _ptr(out), out.stride(0), out.stride(1),
                                           out.stride(2), _stream(X3)), 'mvb_dense_predict_batched')
    return out
