"""Developer timing of the ReceptiveField estimator at a realistic size.
  GPU box:          python tools/rf_bench.py ours
  build container:  MVPY_COMPILE_TORCH=0 python tools/rf_bench.py reference     (unmodified reference on CPU torch)
Same seeded inputs on both sides; prints seconds per fit and, for `ours`, the difference to the reference's
coefficients when profiles/r01_rf_reference_coef.pt exists."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

which = sys.argv[1] if len(sys.argv) > 1 else 'ours'
N, C, F, T = 120, 8, 64, 2000
g = torch.Generator().manual_seed(7)
X = torch.randn(N, C, T, generator=g)
X = X + 0.7 * torch.roll(X, 1, -1)
w = torch.randn(F, C, 61, generator=g) * torch.hann_window(61)
y = torch.zeros(N, F, T)
for k in range(61):
    lag = -10 + k
    Z = torch.zeros_like(X)
    if lag >= 0: Z[..., lag:] = X[..., :T - lag]
    else: Z[..., :T + lag] = X[..., -lag:]
    y += torch.einsum('fc,nct->nft', w[:, :, k], Z)
y = y + 2.0 * y.std() * torch.randn(N, F, T, generator=g)
alphas = torch.logspace(0, 6, 7)
kw = dict(alpha=alphas, reg_cv=5)
if which == 'reference':
    sys.path.insert(0, '/root/reference')
    import warnings; warnings.filterwarnings('ignore')
    from mvpy.estimators import ReceptiveField
    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter()
    m = ReceptiveField(-0.1, 0.5, 100, **kw).fit(X, y)
    dt = time.perf_counter() - t0
    print(f'reference (CPU, {os.cpu_count()} threads): {dt:.2f} s per fit; alpha_ = {float(m.alpha_)}  p = {C * 61}')
    torch.save(dict(coef=m.coef_, alpha=float(m.alpha_), seconds=dt, threads=os.cpu_count()), 'profiles/r01_rf_reference_coef.pt')
else:
    import mvpy_b200 as mv
    from mvpy_b200 import _lib
    Xd, yd = X.cuda(), y.cuda()
    for rep in range(3):
        n0 = _lib.launch_count()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m = mv.estimators.ReceptiveField(-0.1, 0.5, 100, **kw).fit(Xd, yd)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f'ours (B200): {dt * 1e3:.1f} ms per fit; alpha_ = {float(m.alpha_)}  launches = {_lib.launch_count() - n0}', flush=True)
    ref = 'profiles/r01_rf_reference_coef.pt'
    if os.path.exists(ref):
        r = torch.load(ref)
        err = float((m.coef_.cpu() - r['coef']).abs().max() / r['coef'].abs().max())
        print(f'vs reference: alpha {r["alpha"]} (ours {float(m.alpha_)}), max coef error / max |coef| = {err:.2e}; reference took {r["seconds"]:.1f} s on {r["threads"]} threads -> speed-up {r["seconds"] / dt:.0f}x')
