"""mvb_score on a 391 MB problem (ncu target / timing).  python tools/score_probe.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvpy_b200 import _lib
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
R, K = 24, 16 * 306 * 400
y = torch.randn(R, K, device='cuda'); yh = y + 0.1 * torch.randn_like(y); yb = y.mean(0).contiguous()
for _ in range(reps):
    r2, pr = _lib.score(y, yh, yb, True, True)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); _lib.score(y, yh, yb, True, True); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort()
print(f'score {R} x {K}: median {ts[5] * 1e3:.1f} us  {(2 * R * K + 3 * K) * 4 / ts[5] / 1e6:.0f} GB/s')
