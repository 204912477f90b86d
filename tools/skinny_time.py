"""Developer timing of the skinny Lanczos-shaped product (128 x 12928 x 12928, split-K 5) on the TMEM-A kernel: plane format, cache
hints, and the MVB_TC_DBG experiments (3: A values re-read from stage 0 -> no A latency; 4: B copies re-read from stage 0 -> no HBM stream)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvpy_b200 import _lib
lib = _lib.load()
fn = lib.mvb_dev_skinny
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
torch.manual_seed(0)
N = K = 12928
A = torch.randn(128, K, device='cuda'); B = torch.randn(N, K, device='cuda')
rowK = (torch.arange(N, device='cuda') * K).contiguous(); iota = torch.arange(K, device='cuda')
planes = torch.empty(((N + 255) // 256) * ((K + 31) // 32) * 2 * 32896, dtype=torch.uint8, device='cuda')
st = torch.cuda.current_stream().cuda_stream
ref = (A.double() @ B.double().T)
for splits in (5, 3):
    part = torch.zeros(splits, 128, N, device='cuda')
    for mode in (0, 2, 1, 3):
        assert fn(A.data_ptr(), B.data_ptr(), N, K, splits, mode, planes.data_ptr(), part.data_ptr(), rowK.data_ptr(), iota.data_ptr(), 1, st) == 0
        for _ in range(3):
            fn(A.data_ptr(), B.data_ptr(), N, K, splits, mode, planes.data_ptr(), part.data_ptr(), rowK.data_ptr(), iota.data_ptr(), 0, st)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(A.data_ptr(), B.data_ptr(), N, K, splits, mode, planes.data_ptr(), part.data_ptr(), rowK.data_ptr(), iota.data_ptr(), 0, st)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        err = ((part.sum(0).double() - ref).norm() / ref.norm()).item()
        t = sorted(ts)[len(ts) // 2]
        print(f'dbg {os.environ.get("MVB_TC_DBG", "0")} splits {splits} mode {mode} ({"raw" if mode & 1 else "hi/lo"} planes{", hints" if mode & 2 else ""}): '
              f'{t * 1e3:7.1f} us  {2 * 128 * N * K / t * 1e-9:6.1f} TFLOP/s  rel err {err:.2e}', flush=True)
