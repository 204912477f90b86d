"""Developer probe: mvb_dev_chol_inv on 500 x 20 problems of 306 (padded 320) columns -- total time, per-phase cycles of block 0, and
the residual |L^-1 (G + alpha I) L^-T - I| of a few problems."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvpy_b200 import _lib
lib = _lib.load()
fn = lib.mvb_dev_chol_inv
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 2 + [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 4
torch.manual_seed(0)
B, p, A, n = 500, int(sys.argv[1]) if len(sys.argv) > 1 else 306, 20, 1600
Pp = (p + 15) // 16 * 16
tri_rows = (Pp + 127) // 128 * 128
X = torch.randn(B, n, Pp, device='cuda')
X[:, :, p:] = 0
G = torch.bmm(X.transpose(1, 2), X).contiguous()
al = torch.logspace(-5, 5, A).cuda()
Li = torch.empty(B * A, tri_rows, Pp, device='cuda')
status = torch.zeros(B, dtype=torch.int32, device='cuda')
clk = torch.zeros(12, dtype=torch.int64, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for it in range(3):
    clk.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(G.data_ptr(), p, B, al.data_ptr(), A, Li.data_ptr(), status.data_ptr(), clk.data_ptr() if it == 2 else None, st)
    e1.record(); torch.cuda.synchronize()
    assert rc == 0, lib.mvb_last_error()
    print(f'chol_inv {B * A} problems p={p}: {e0.elapsed_time(e1):.2f} ms')
names = ['load', 'diag factor (rest: sync)', 'panel solve', 'trailing update', 'diag inverse', 'panel product', 'inverse update', 'store', '  diag: load', '  diag: factor', '  diag: write-back', '-']
c = clk.cpu().tolist()
for nm, v in zip(names, c):
    print(f'  {nm:16s} {v:9d} cycles  {100.0 * v / max(1, sum(c)):5.1f} %')
for prob in (0, 7 * A + 3, B * A - 1):
    b, a = divmod(prob, A)
    M = G[b, :p, :p].double() + al[a].double() * torch.eye(p, device='cuda', dtype=torch.float64)
    L = Li[prob, :p, :p].double()
    R = L @ M @ L.T - torch.eye(p, device='cuda', dtype=torch.float64)
    print(f'  problem {prob}: |Li M Li^T - I|_max = {R.abs().max().item():.2e}, upper triangle max {Li[prob, :Pp].triu(1).abs().max().item():.1e}, status {int(status[b])}')
