import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvpy_b200 import _lib
torch.manual_seed(0)
X = torch.randn(1600, 306, 500, device='cuda'); y = torch.randn(1600, 3, 500, device='cuda')
al = torch.logspace(-5, 5, 20).cuda()
for _ in range(2):
    r = _lib.dense_ridge_batched(X, y, None, al, True, True, want_gram=True)
torch.cuda.synchronize()
print('ok', int(r['status'].sum()))
