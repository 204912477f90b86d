"""Developer probe: per-slice host (enqueue) and total time of Sliding(RidgeDecoder).fit at BASELINE configs[4] slice size,\nplus a bare decoder / RidgeCV fit (run on a GPU box):  [MVB_SLIDING_LANES=n] [MVB_SLIDING_GRAPHS=0] python tools/cfg5_probe.py"""
import os, sys, time
sys.path.insert(0, '/root/repo')
import torch
import mvpy_b200 as mv
from mvpy_b200 import _lib
T = 128
torch.manual_seed(0)
N, C, F = 2000, 306, 3
X = torch.randn(N, C, T, device='cuda')
B = torch.randn(C, F, device='cuda') * 0.05
y = torch.einsum('nct,cf->nft', X, B) + torch.randn(N, F, T, device='cuda')
alphas = torch.logspace(-5, 5, 20)
Xtr, ytr = X[:1600].contiguous(), y[:1600].contiguous()
for rep in range(3):
    model = mv.estimators.Sliding(mv.estimators.RidgeDecoder(alphas=alphas), dims=-1)
    n0 = _lib.launch_count()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    model.fit(Xtr, ytr)
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'fit {T} slices: enqueue {1e3*(t1-t0)/T:.2f} ms/slice, total {1e3*(t2-t0)/T:.2f} ms/slice, launches/slice {( _lib.launch_count()-n0)/T:.0f}', flush=True)
# one plain decoder fit, timed in pieces
d = mv.estimators.RidgeDecoder(alphas=alphas)
Xs, ys = Xtr[..., 0].contiguous(), ytr[..., 0].contiguous()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): d.fit(Xs, ys)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'decoder.fit: enqueue {1e3*(t1-t0)/20:.2f} ms, total {1e3*(t2-t0)/20:.2f} ms', flush=True)
r = mv.estimators.RidgeCV(alphas=alphas, alpha_per_target=True)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): r.fit(Xs, ys)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'ridgecv.fit: enqueue {1e3*(t1-t0)/20:.2f} ms, total {1e3*(t2-t0)/20:.2f} ms', flush=True)
