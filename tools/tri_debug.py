import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvpy_b200 as mv
from mvpy_b200 import _lib
n, p, f, t = 420, 320, 1, 5
g = torch.Generator().manual_seed(100 + p)
X = torch.randn(n, p, t, generator=g)
y = torch.einsum('npt,pf->nft', X, torch.randn(p, f, generator=g) * 0.1) + torch.randn(n, f, t, generator=g)
X = X + 3.0 * torch.randn(1, p, 1, generator=g)
al = torch.logspace(-3, 4, 8)
r = _lib.dense_ridge_batched(X.cuda(), y.cuda(), None, al.cuda(), True, True, want_gram=True)
print('status', r['status'].cpu().tolist(), 'best', r['best'].cpu().reshape(-1).tolist())
n0 = _lib.launch_count()
sl = mv.estimators.Sliding(mv.estimators.RidgeDecoder(alphas=al), dims=-1).fit(X.cuda(), y.cuda())
print('launches', _lib.launch_count() - n0)
for mk in (lambda: mv.estimators.RidgeCV(alphas=al, alpha_per_target=False), lambda: mv.estimators.RidgeCV(alphas=al, fit_intercept=False, alpha_per_target=True)):
    n0 = _lib.launch_count()
    sl = mv.estimators.Sliding(mk(), dims=-1).fit(X.cuda(), y.cuda())
    print('launches', _lib.launch_count() - n0)
r = _lib.dense_ridge_batched(X.cuda(), y.cuda(), None, al.cuda(), False, True)
print('no intercept: status', r['status'].cpu().tolist(), 'best', r['best'].cpu().reshape(-1).tolist(), r['loss'][0, :, 0].cpu().tolist())
