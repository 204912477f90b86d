"""Developer check of the band path (block-Lanczos reduction + block-Cholesky LOO sweep) against fp64 (run on a GPU box).

    python tools/band_check.py [C ...]       # stimulus features per case (p = C * 201), default 3 8

For each case: orthogonality of Qt, reconstruction  Qt^T T Qt - pad(G),  then mvb_ridge_solve (band) vs the fp64 oracle
(loss, coefficients, chosen alpha) and the time of the reduction / the whole solve.  Knobs come from the environment
(MVB_LB_POLISH_NS, MVB_LB_EIG_TOL, MVB_LB_EIG_SWEEPS, MVB_LB_REORTH3, MVB_BAND_MIN_P).
"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mvpy_b200 import _lib as L
from mvpy_b200.dataset import make_meeg_continuous

dev = torch.device('cuda')


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def dense_T(Td, Ts):
    nb = Td.shape[0]
    P = nb * 128
    T = torch.zeros(P, P, dtype=torch.float64, device=Td.device)
    for j in range(nb):
        T[j * 128:(j + 1) * 128, j * 128:(j + 1) * 128] = Td[j].double()
        if j > 0:
            T[j * 128:(j + 1) * 128, (j - 1) * 128:j * 128] = Ts[j].double()
            T[(j - 1) * 128:j * 128, j * 128:(j + 1) * 128] = Ts[j].double().T
    return T


def run(C, F=64, oracle=True):
    torch.manual_seed(0)
    X, y = make_meeg_continuous(fs=200, n_features=C, n_channels=F)
    X = X[:96].contiguous()
    y = y[:96].contiguous()
    sd = X.std(0, keepdim=True)
    sd[(sd == 0) | sd.isnan()] = 1.0
    X = (X - X.mean(0, keepdim=True)) / sd
    alphas = torch.logspace(-5, 5, 20)
    Xd, yd, ad = X.to(dev), y.to(dev), alphas.to(dev)
    xp = L.trf_pad(Xd, 200, 0)
    d = L.Design(xp, 400, 201)
    G, XtY, mu, ybar = L.trf_stats(d, yd, True)
    p = G.shape[0]
    torch.cuda.synchronize()
    L.band_reduce(G)
    e0 = ev()
    Qt, Td, Ts, cnt = L.band_reduce(G, counters=True)
    e1 = ev()
    torch.cuda.synchronize()
    t_red = e0.elapsed_time(e1)
    P = Qt.shape[0]
    Q64 = Qt.double()
    orth = (Q64 @ Q64.T - torch.eye(P, dtype=torch.float64, device=dev)).abs().max().item()
    Gp = torch.zeros(P, P, dtype=torch.float64, device=dev)
    Gp[:p, :p] = G.double()
    g0 = G.diagonal().mean().item()
    for r in range(p, P):
        Gp[r, r] = abs(g0) * (1.0 + 0.01 * (r - p))
    T = dense_T(Td, Ts)
    rec = (Q64.T @ T @ Q64 - Gp).abs().max().item() / Gp.abs().max().item()
    # off-band energy of the exact projection
    Tex = Q64 @ Gp @ Q64.T
    mask = torch.ones(P // 128, P // 128, device=dev).tril(-2) + torch.ones(P // 128, P // 128, device=dev).triu(2)
    offband = (Tex * mask.repeat_interleave(128, 0).repeat_interleave(128, 1)).abs().max().item() / Gp.abs().max().item()
    msg = f'C={C} p={p} P={P}: reduce {t_red:.1f} ms  orth={orth:.1e} rec={rec:.1e} offband={offband:.1e} fallbacks(chol,ns)={cnt}'
    del Q64, T, Tex, Gp
    print(msg, flush=True)
    msg = '   '
    # full solve
    G2 = G.clone()
    L.ridge_solve(d, yd, G2, XtY, mu, ybar, ad, True, False, want_loss=True)
    G2 = G.clone()
    e0 = ev()
    coef, ic, al, best, loss = L.ridge_solve(d, yd, G2, XtY, mu, ybar, ad, True, False, want_loss=True)
    e1 = ev()
    torch.cuda.synchronize()
    msg += f'  solve {e0.elapsed_time(e1):.1f} ms best={best.cpu().tolist()}'
    if oracle:
        # fp64 reference on the GPU via torch (eigh of the fp64 Gram built from the fp32 inputs)
        from oracle import mvpy_oracle as orc
        win = np.arange(-200, 1)
        Xt = orc.delay_design(X.double(), win).to(dev)
        yt = orc.flatten_targets(y.double()).to(dev)
        Xc = Xt - Xt.mean(0, keepdim=True)
        yc = yt - yt.mean(0, keepdim=True)
        G64 = Xc.T @ Xc
        lam, V = torch.linalg.eigh(G64)
        Z = Xc @ V
        Vtb = V.T @ (Xc.T @ yc)
        losses, coefs = [], []
        n = Xc.shape[0]
        for a in alphas.double().tolist():
            w = 1.0 / (lam + a)
            beta = V @ (w[:, None] * Vtb)
            r = yc - Xc @ beta
            h = (Z * Z) @ w + 1.0 / n
            losses.append(((r / (1 - h)[:, None]) ** 2).mean(0))
            coefs.append(beta.T)
        losses = torch.stack(losses)
        b_ref = int(losses.mean(1).argmin())
        eL = ((loss.double() - losses).abs() / losses.abs()).max().item()
        eC = (coef.double() - coefs[b_ref]).abs().max().item() / coefs[b_ref].abs().max().item()
        eCf = ((coef.double() - coefs[b_ref]).norm() / coefs[b_ref].norm()).item()
        msg += f' (ref {b_ref})  loss_rel={eL:.1e} coef_max={eC:.1e} coef_fro={eCf:.1e} cond={lam.max().item() / lam.min().item():.1e}'
    print(msg, flush=True)


if __name__ == '__main__':
    Cs = [int(a) for a in sys.argv[1:]] or [3, 8]
    print('knobs:', {k: v for k, v in os.environ.items() if k.startswith('MVB_')}, flush=True)
    for C in Cs:
        run(C)
