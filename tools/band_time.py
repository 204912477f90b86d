"""Developer timing: the block-Lanczos reduction alone on a cfg2-sized random SPD matrix (p = 12864); MVB_TC_DBG experiments give invalid results."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvpy_b200 import _lib
torch.manual_seed(0)
p = 12864
A = torch.randn(p, 2 * p, device='cuda')
G = (A @ A.T).contiguous()
del A
_lib.band_reduce(G.clone())
torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    Gc = G.clone()
    e0.record(); _lib.band_reduce(Gc); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print('knobs', {k: v for k, v in os.environ.items() if k.startswith('MVB_')}, 'band reduce ms', [round(t, 1) for t in ts])
