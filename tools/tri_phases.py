"""Developer probe: per-phase cycles of block (0, 0) of rows_apply_leverage_kernel in one batched configs[4] call."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvpy_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
X = torch.randn(1600, 306, 500, device='cuda'); y = torch.randn(1600, 3, 500, device='cuda')
al = torch.logspace(-5, 5, 20).cuda()
for _ in range(2):
    _lib.dense_ridge_batched(X, y, None, al, True, True, want_gram=True)
out = (ctypes.c_longlong * 12)()
assert lib.mvb_dev_rows_apply_phases(out) == 0
for nm, v in zip(['rows: load tile + factors', 'rows: reflectors', 'rows: write-back', 'rows: leverages', 'tridiag: column + norm', 'tridiag: v, scalars', 'tridiag: product pass', 'tridiag: combine', 'tridiag: dot + w', 'tridiag: rank-2 update', '-', '-'], out):
    print(f'{nm:22s} {v:9d} cycles')
