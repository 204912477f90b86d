"""Developer timing of the fit stages on one fold-sized problem (run on a GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvpy_b200 as mv
from mvpy_b200 import _lib as L

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

def run(C, F, N=96, T=400, lag=200, A=20):
    torch.manual_seed(0)
    X = torch.randn(N, C, T, device='cuda'); y = torch.randn(N, F, T, device='cuda')
    al = torch.logspace(-5, 5, A).cuda()
    for rep in range(2):
        t = [ev()]
        xp = L.trf_pad(X, lag, 0); d = L.Design(xp, T, lag + 1); t.append(ev())
        G, XtY, mu, ybar = L.trf_stats(d, y, True); t.append(ev())
        G2 = G.clone(); t.append(ev())
        Vt, lam = L.eigh(G2[None]); t.append(ev())
        coef, ic, a, best, loss = L.ridge_solve(d, y, G, XtY, mu, ybar, al, True, False); t.append(ev())
        yh = L.trf_predict(d, coef, ic); t.append(ev())
        r2, pr = L.score(y.reshape(N, -1), yh.reshape(N, -1), None, True, True); t.append(ev())
        torch.cuda.synchronize()
    names = ['pad', 'stats(gram)', 'clone', 'eigh alone', 'solve(incl eigh)', 'predict', 'score']
    print(f'C={C} F={F} p={C*(lag+1)} n={N*T}: ' + '  '.join(f'{n}={t[i].elapsed_time(t[i+1]):.2f}ms' for i, n in enumerate(names)))

run(3, 64)
run(8, 306)
if len(sys.argv) > 1:
    run(16, 306)
