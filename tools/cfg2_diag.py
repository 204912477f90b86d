"""Diagnostic: cfg2 fold 0, ours vs fp64 oracle pieces (run on a GPU box).  python tools/cfg2_diag.py [F] [C] [reduce]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvpy_b200 as mv
from mvpy_b200 import _lib
from oracle import mvpy_oracle as orc

F = int(sys.argv[1]) if len(sys.argv) > 1 else 306
C = int(sys.argv[2]) if len(sys.argv) > 2 else 64
print('knobs:', {k: v for k, v in os.environ.items() if k.startswith('MVB_')}, 'F', F, 'C', C, flush=True)
torch.manual_seed(0)
X, y = mv.dataset.make_meeg_continuous(fs=200, n_features=C, n_channels=F)
alphas = torch.logspace(-5, 5, 20)
tr, te = orc.kfold_indices(120, 5)[0]
sc = mv.preprocessing.Scaler().to_torch().fit(X[tr].cuda())
Xs = sc.transform(X[tr].cuda())
td = mv.estimators.TimeDelayed(-1.0, 0.0, 200, alphas=alphas).fit(Xs, y[tr].cuda())
design = td._design(Xs)
G, XtY, mu, ybar = _lib.trf_stats(design, y[tr].cuda().contiguous(), True)
win = orc.lag_window(-1.0, 0.0, 200)
Xt = orc.delay_design(Xs.double(), win)
yt = orc.flatten_targets(y[tr].cuda().double())
Xc = Xt - Xt.mean(0, keepdim=True)
yc = yt - yt.mean(0, keepdim=True)
G64 = Xc.T @ Xc
b64 = Xc.T @ yc
print('G rel', ((G.double() - G64).norm() / G64.norm()).item(), 'XtY rel', ((XtY.double() - b64).norm() / b64.norm()).item())
fit = orc.ridge_loo_fit_chol(Xt, yt, alphas.cuda().double(), True, False)
L = td.estimator.loss_.double()
ours = td.estimator.coef_.double()
print('alpha ours', float(td.estimator.alpha_), 'oracle', float(fit['alpha']))
print('loss rel max per alpha', ((L - fit['losses']).abs() / fit['losses'].abs()).max(1).values.cpu().numpy())
print('coef rel', ((ours - fit['coef']).norm() / fit['coef'].norm()).item())
# coefficient error at the chosen alpha if the solve used the exact fp64 Gram of the same inputs but ours' XtY etc.
a = float(td.estimator.alpha_)
eye = torch.eye(G64.shape[0], dtype=torch.float64, device='cuda')
beta_mixed = torch.linalg.solve(G.double() + a * eye, XtY.double())
print('coef: fp64 solve on OUR fp32 G / XtY vs oracle', ((beta_mixed.T - fit['coef']).norm() / fit['coef'].norm()).item(),
      '; ours vs that', ((ours - beta_mixed.T).norm() / beta_mixed.norm()).item())
if len(sys.argv) > 3:
    _lib.band_reduce(G.clone())
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    Qt, Td, Ts, cnt = _lib.band_reduce(G.clone(), counters=True)
    e1.record(); torch.cuda.synchronize()
    print('band reduce ms', e0.elapsed_time(e1))
    P = Qt.shape[0]
    Q64 = Qt.double()
    print('orth', (Q64 @ Q64.T - torch.eye(P, dtype=torch.float64, device='cuda')).abs().max().item(), 'fallbacks', cnt)
