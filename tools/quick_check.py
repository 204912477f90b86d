"""Developer smoke check of the C-ABI stages against the CPU oracle (run on a GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mvpy_b200 import _lib as L
from oracle import mvpy_oracle as orc

torch.manual_seed(0)
dev = torch.device('cuda')

def run(dtype, N=12, C=3, T=60, F=5, lag_neg=7, lag_pos=2, A=9, per_target=False, icpt=True):
    X = torch.randn(N, C, T, dtype=dtype); B = torch.randn(F, C, dtype=dtype) * 0.5
    y = torch.einsum('fc,nct->nft', B, X) + torch.randn(N, F, T, dtype=dtype) + 0.3
    X = X + 0.2
    win = np.arange(-lag_neg, lag_pos + 1)
    W = len(win)
    alphas = torch.logspace(-2, 4, A, dtype=dtype)
    Xd, yd, ad = X.to(dev), y.to(dev).contiguous(), alphas.to(dev)
    xp = L.trf_pad(Xd, lag_neg, lag_pos)
    d = L.Design(xp, T, W)
    G, XtY, mu, ybar = L.trf_stats(d, yd, icpt)
    # oracle
    Xt, yt = orc.delay_design(X.double(), win), orc.flatten_targets(y.double())
    ref = orc.ridge_loo_fit_primal(Xt, yt, alphas.double(), icpt, per_target)
    eG = (G.cpu().double() - ref['G']).abs().max().item() / ref['G'].abs().max().item()
    eb = (XtY.cpu().double() - ref['b']).abs().max().item() / ref['b'].abs().max().item()
    Gc = G.clone()
    Vt, lam = L.eigh(Gc[None].clone())
    rec = (Vt[0].T * lam[0]) @ Vt[0]
    eE = (rec - G).abs().max().item() / G.abs().max().item()
    eO = (Vt[0] @ Vt[0].T - torch.eye(Vt.shape[1], device=dev, dtype=dtype)).abs().max().item()
    coef, ic, al, best, loss = L.ridge_solve(d, yd, G, XtY, mu, ybar, ad, icpt, per_target, want_loss=True)
    eL = ((loss.cpu().double() - ref['losses']).abs() / ref['losses'].abs()).max().item()
    eC = (coef.cpu().double() - ref['coef']).abs().max().item() / ref['coef'].abs().max().item()
    same_alpha = torch.equal(best.cpu().long().reshape(-1), ref['best'].reshape(-1))
    yh = L.trf_predict(d, coef, ic if icpt else None)
    yh_ref = orc.trf_predict(X.double(), win, ref['coef'], ref['intercept'])
    eP = (yh.cpu().double() - yh_ref).abs().max().item()
    # scoring
    R, K = N, F * T
    r2, pr = L.score(yd.reshape(R, K), yh.reshape(R, K), None, True, True)
    r2_ref = orc.r2_last(orc.move_reduce_last(y.double(), (0,)), orc.move_reduce_last(yh_ref, (0,)))
    pr_ref = orc.pearsonr_last(orc.move_reduce_last(y.double(), (0,)), orc.move_reduce_last(yh_ref, (0,)))
    eR = (r2.cpu().double().reshape(F, T) - r2_ref).abs().max().item()
    eq = (pr.cpu().double().reshape(F, T) - pr_ref).abs().max().item()
    print(f'{str(dtype):14s} p={C*W:4d} n={N*T:5d} G={eG:.1e} XtY={eb:.1e} eig_rec={eE:.1e} orth={eO:.1e} loss={eL:.1e} coef={eC:.1e} '
          f'alpha_ok={same_alpha} pred={eP:.1e} r2={eR:.1e} pearson={eq:.1e}')

for dt in (torch.float64, torch.float32):
    run(dt)
    run(dt, per_target=True)
    run(dt, icpt=False)
    run(dt, N=30, C=4, T=100, F=7, lag_neg=40, lag_pos=0, A=20)
t0 = time.time()
run(torch.float32, N=96, C=3, T=400, F=64, lag_neg=200, lag_pos=0, A=20)
torch.cuda.synchronize(); print('cfg1-size fold incl. oracle', time.time() - t0)
print('launches', L.launch_count())
