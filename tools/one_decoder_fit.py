"""One RidgeDecoder fit at BASELINE configs[4] slice size between cudaProfilerStart/Stop, for ncu launch lists."""
import sys
sys.path.insert(0, '/root/repo')
import torch
import mvpy_b200 as mv
torch.manual_seed(0)
X = torch.randn(1600, 306, device='cuda'); y = torch.randn(1600, 3, device='cuda') + X[:, :3]
al = torch.logspace(-5, 5, 20)
d = mv.estimators.RidgeDecoder(alphas=al)
d.fit(X, y); torch.cuda.synchronize()
torch.cuda.profiler.start()
d.fit(X, y); torch.cuda.synchronize()
torch.cuda.profiler.stop()
