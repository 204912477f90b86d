"""Developer timing of one dense RidgeCV fit (the Sliding/RidgeDecoder unit of BASELINE configs[4]: n=1600, p=306) through
both solver routes (run on a GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvpy_b200 as mv

def run(n, p, F, route, reps=10):
    os.environ['MVB_BAND_MIN_P'] = '128' if route == 'band' else '1000000'
    torch.manual_seed(0)
    X = torch.randn(n, p, device='cuda'); B = torch.randn(p, F, device='cuda') * 0.1
    y = X @ B + torch.randn(n, F, device='cuda')
    al = torch.logspace(-5, 5, 20)
    m = mv.estimators.RidgeCV(alphas=al, alpha_per_target=True)
    m.fit(X, y); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        m.fit(X, y)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return dt, m.alpha_.cpu(), m.coef_.cpu()

for (n, p, F) in [(1600, 306, 3), (1600, 160, 3), (4000, 480, 8), (1600, 306, 64)]:
    te, ae, ce = run(n, p, F, 'eigen')
    tb, ab, cb = run(n, p, F, 'band')
    print(f'n={n} p={p} F={F}: eigen {te*1e3:.2f} ms  band {tb*1e3:.2f} ms  alpha_equal={torch.equal(ae, ab)}  coef_rel={((ce-cb).norm()/ce.norm()).item():.1e}', flush=True)
