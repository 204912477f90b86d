import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvpy_b200 import _lib as L
for p in [int(a) for a in sys.argv[1:]] or [200, 603, 1608]:
    g = torch.Generator().manual_seed(p)
    A = torch.randn(p + 300, p, generator=g, dtype=torch.float64) * torch.logspace(0, -2, p, dtype=torch.float64)
    G = (A.T @ A).float().cuda()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        Vt, lam = L.eigh(G[None].clone())
        torch.cuda.synchronize(); dt = time.time() - t0
    Vt, lam = Vt[0].double(), lam[0].double()
    Gd = G.double()
    rec = ((Vt.T * lam) @ Vt - Gd).abs().max().item() / Gd.abs().max().item()
    orth = (Vt @ Vt.T - torch.eye(p, device='cuda', dtype=torch.float64)).abs().max().item()
    ref = torch.linalg.eigvalsh(Gd)
    ev = (lam.sort().values - ref).abs().max().item() / ref.max().item()
    t0 = time.time(); torch.linalg.eigh(G); torch.cuda.synchronize(); t1 = time.time() - t0
    t0 = time.time(); torch.linalg.eigh(G); torch.cuda.synchronize(); t1 = time.time() - t0
    print(f'p={p}: {dt*1e3:.1f} ms  recon={rec:.2e} orth={orth:.2e} eigval_err={ev:.2e}   (torch.linalg.eigh fp32: {t1*1e3:.1f} ms)', flush=True)
