"""Developer timing of BASELINE configs[4] (Sliding RidgeDecoder, 2000 trials x 306 channels, 20 alphas, 5 folds) on a
subset of the 500 timepoints (run on a GPU box):  python tools/cfg5_check.py [n_timepoints]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvpy_b200 as mv
from mvpy_b200 import _lib

T = int(sys.argv[1]) if len(sys.argv) > 1 else 50
torch.manual_seed(0)
N, C, F = 2000, 306, 3
X = torch.randn(N, C, T, device='cuda')
B = torch.randn(C, F, device='cuda') * 0.05
hann = torch.hann_window(T, device='cuda') if T > 2 else torch.ones(T, device='cuda')
y = torch.einsum('nct,cf->nft', X, B) * hann + torch.randn(N, F, T, device='cuda')
alphas = torch.logspace(-5, 5, 20)
model = mv.estimators.Sliding(mv.estimators.RidgeDecoder(alphas=alphas), dims=-1)
metric = (mv.metrics.r2, mv.metrics.pearsonr)
for rep in range(2):
    n0 = _lib.launch_count()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    val, sc = mv.crossvalidation.cross_val_score(model, X, y, cv=5, metric=metric)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    fits = T * 5 * 20
    print(f'T={T}: {dt:.2f} s  {fits / dt:.0f} fits/s  ({dt / (T * 5) * 1e3:.2f} ms per slice-fit incl. predict+score)  launches={_lib.launch_count() - n0} '
          f'r2 mean={float(sc["r2"].mean()):.4f}', flush=True)
