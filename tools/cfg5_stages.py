"""Stage timing of one cfg5 fold (Sliding RidgeDecoder, 1600 x 306 x 500 -> 3 targets): where the time outside the kernels goes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvpy_b200 as mv
from mvpy_b200 import _lib, _batched

torch.manual_seed(0)
N, C, T, F = 2000, 306, 500, 3
X = torch.randn(N, C, T, device='cuda')
y = torch.randn(N, F, T, device='cuda')
al = torch.logspace(-5, 5, 20)
train = torch.arange(400, 2000, device='cuda')
test = torch.arange(0, 400, device='cuda')

def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    print(f'{name:50s} {(time.perf_counter() - t0) / reps * 1e3:8.2f} ms', flush=True)
    return out

Xtr = timed('X[train], y[train] (ATen gather)', lambda: (X[train], y[train]))
Xt, yt = Xtr
res = timed('dense_ridge_batched (kernels only)', lambda: _lib.dense_ridge_batched(Xt, yt, None, al.cuda(), True, True, want_gram=True))
timed('dense_ridge_batched via sample_idx on full X', lambda: _lib.dense_ridge_batched(X, y, train, al.cuda(), True, True, want_gram=True))
timed('patterns', lambda: _batched._patterns(res['gram'], res['coef'], yt, 1600))
dec = mv.estimators.RidgeDecoder(alphas=al)
pairs = [(i, i) for i in range(T)]
fitted = timed('_batched.fit_slices (kernels + 500 estimator objects)', lambda: _batched.fit_slices(dec, Xt, yt, pairs))
sl = mv.estimators.Sliding(dec, dims=-1)
timed('Sliding.fit', lambda: sl.fit(Xt, yt))
Xte, yte = X[test], y[test]
yh = timed('Sliding.predict (batched)', lambda: sl.predict(Xte))
m = (mv.metrics.r2, mv.metrics.pearsonr)
timed('metrics.score r2 + pearsonr', lambda: mv.metrics.score(sl, m, Xte, yte))
from mvpy_b200.crossvalidation.validator import fit_model_, _move_tensors
import copy
mm = copy.deepcopy(m)
r = timed('fit_model_ (one fold, two metrics)', lambda: fit_model_(sl, train, test, X, y=y, metric=mm), reps=2)
timed('_move_tensors(model) to the input device (cuda)', lambda: _move_tensors(r['model'], X.device), reps=2)
timed('cross_val_score 5 folds', lambda: mv.crossvalidation.cross_val_score(sl, X, y, cv=5, metric=m), reps=1)
# per-kernel times of the batched call (events around each stage are not exposed: use the profiler hook classes)
_lib.profile_reset()
_lib.dense_ridge_batched(Xt, yt, None, al.cuda(), True, True, want_gram=True)
torch.cuda.synchronize()
for k, v in _lib.profile_read().items():
    print(f'  {k[:40]:40s} {v["ms"]:8.2f} ms  {v["flops"] / v["ms"] * 1e-9:7.1f} TFLOP/s')
