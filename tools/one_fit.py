"""One TimeDelayed fit (+predict+score) at a named size, for ncu launch lists / captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvpy_b200 as mv
cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
C, F = (64, 306) if cfg == 'cfg2' else (3, 64)
torch.manual_seed(0)
X = torch.randn(96, C, 400, device='cuda'); y = torch.randn(96, F, 400, device='cuda')
td = mv.estimators.TimeDelayed(-1.0, 0.0, 200, alphas=torch.logspace(-5, 5, 20))
for _ in range(reps):
    td.fit(X, y)
    yh = td.predict(X[:24])
    r = mv.math.r2(yh.movedim(0, -1), y[:24].movedim(0, -1))
torch.cuda.synchronize()
print('ok', cfg, float(td.estimator.alpha_), float(r.mean()))
