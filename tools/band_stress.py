"""Developer stress check of the band route on ill-conditioned designs (strongly autocorrelated / collinear stimuli),
against an fp64 solve of the same normal equations (run on a GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mvpy_b200 import _lib as L
dev = torch.device('cuda')

def ar1(n, c, t, phi, g):
    e = torch.randn(n, c, t, generator=g)
    x = torch.zeros_like(e)
    x[..., 0] = e[..., 0]
    for i in range(1, t):
        x[..., i] = phi * x[..., i - 1] + (1 - phi * phi) ** 0.5 * e[..., i]
    return x

def run(name, X, F=16, lags=120):
    g = torch.Generator().manual_seed(3)
    N, C, T = X.shape
    B = torch.randn(F, C, generator=g) * 0.5
    y = torch.einsum('fc,nct->nft', B, X) + torch.randn(N, F, T, generator=g)
    alphas = torch.logspace(-5, 5, 20)
    Xd, yd, ad = X.to(dev), y.to(dev).contiguous(), alphas.to(dev)
    xp = L.trf_pad(Xd, lags, 0)
    d = L.Design(xp, T, lags + 1)
    G, XtY, mu, ybar = L.trf_stats(d, yd, True)
    p = G.shape[0]
    Qt, Td, Ts, cnt = L.band_reduce(G, counters=True)
    P = Qt.shape[0]
    Q64 = Qt.double()
    orth = (Q64 @ Q64.T - torch.eye(P, dtype=torch.float64, device=dev)).abs().max().item()
    coef, ic, al, best, loss = L.ridge_solve(d, yd, G.clone(), XtY, mu, ybar, ad, True, False, want_loss=True)
    # fp64 reference from the fp64 design
    t_idx = (torch.arange(T)[None, :] + torch.arange(-lags, 1)[:, None]).clamp_(0, T - 1)
    Xt = X.double()[:, :, t_idx].permute(0, 3, 1, 2).reshape(N * T, -1).to(dev)
    yt = y.double().permute(0, 2, 1).reshape(N * T, F).to(dev)
    Xc, yc = Xt - Xt.mean(0, keepdim=True), yt - yt.mean(0, keepdim=True)
    G64 = Xc.T @ Xc
    lam, V = torch.linalg.eigh(G64)
    Z = Xc @ V
    Vtb = V.T @ (Xc.T @ yc)
    n = Xc.shape[0]
    losses, coefs = [], []
    for a in alphas.double().tolist():
        w = 1.0 / (lam + a)
        beta = V @ (w[:, None] * Vtb)
        r = yc - Xc @ beta
        h = (Z * Z) @ w + 1.0 / n
        losses.append(((r / (1 - h)[:, None]) ** 2).mean(0)); coefs.append(beta.T)
    losses = torch.stack(losses)
    b_ref = int(losses.mean(1).argmin())
    eL = ((loss.double() - losses).abs() / losses.abs()).max().item()
    # prediction-space error is what scores see: || X (coef - ref) || / || X ref ||
    dpred = (Xc @ (coef.double() - coefs[b_ref]).T).norm() / (Xc @ coefs[b_ref].T).norm()
    print(f'{name}: p={p} cond(G)={lam.max().item() / max(lam.min().item(), 1e-300):.1e} orth={orth:.1e} fallbacks={cnt} '
          f'best={best.item()} ref={b_ref} loss_rel={eL:.1e} coef_fro={((coef.double() - coefs[b_ref]).norm() / coefs[b_ref].norm()).item():.1e} '
          f'pred_rel={dpred.item():.1e} finite={bool(torch.isfinite(coef).all())}', flush=True)

g = torch.Generator().manual_seed(0)
run('white', torch.randn(40, 4, 300, generator=g))
run('ar1 0.9', ar1(40, 4, 300, 0.9, g))
run('ar1 0.99', ar1(40, 4, 300, 0.99, g))
run('ar1 0.999', ar1(40, 4, 300, 0.999, g))
X = ar1(40, 4, 300, 0.9, g); X[:, 3] = X[:, 2] + 1e-3 * torch.randn(40, 300, generator=g)
run('collinear 1e-3', X)
X = ar1(40, 4, 300, 0.9, g); X[:, 3] = X[:, 2]
run('duplicate feature', X)

def debug(X, F=4, lags=120):
    g = torch.Generator().manual_seed(3)
    N, C, T = X.shape
    y = torch.randn(N, F, T, generator=g)
    alphas = torch.logspace(-5, 5, 20)
    Xd, yd, ad = X.to(dev), y.to(dev).contiguous(), alphas.to(dev)
    xp = L.trf_pad(Xd, lags, 0)
    d = L.Design(xp, T, lags + 1)
    G, XtY, mu, ybar = L.trf_stats(d, yd, True)
    Qt, Td, Ts = L.band_reduce(G)
    print('T finite', bool(torch.isfinite(Td).all()), bool(torch.isfinite(Ts).all()), 'Tdiag max', Td.abs().max().item(), 'G max', G.abs().max().item())
    coef, ic, al, best, loss = L.ridge_solve(d, yd, G.clone(), XtY, mu, ybar, ad, True, False, want_loss=True)
    print('loss[:,0]', loss[:, 0].cpu().tolist())
    print('coef finite', bool(torch.isfinite(coef).all()))

if len(sys.argv) > 1:
    X = ar1(40, 4, 300, 0.9, g); X[:, 3] = X[:, 2]
    debug(X)
