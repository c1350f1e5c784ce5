"""GPU-box probe: cost of the serial phases of potrf_lnlike_kernel (n = 64: one diagonal block only;
n = 128/192: plus one/two TRSM tiles), single CTA vs full grid."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import starry_process_b200 as spb
from starry_process_b200 import _lib
dev = torch.device("cuda:0")
c = spb.get_context(0); lib, ctx = c.lib, c.handle
P = lambda x: ctypes.c_void_p(x.data_ptr())
def timeit(B, n, M, reps=5):
    A = torch.randn(B, n, 32, dtype=torch.float64, device=dev)
    K0 = torch.bmm(A, A.transpose(1, 2)) / 32 + torch.eye(n, dtype=torch.float64, device=dev)
    R0 = torch.randn(B, max(M, 1), n, dtype=torch.float64, device=dev)
    ll = torch.zeros(B, dtype=torch.float64, device=dev); info = torch.zeros(B, dtype=torch.int32, device=dev)
    best = 1e9
    for r in range(reps + 1):
        K = K0.clone(); R = R0.clone(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.spb_cholesky_lnlike(ctx, B, n, P(K), n, n * n, M, P(R), n, max(M, 1) * n, P(ll), None, None, P(info), None))
        e1.record(); torch.cuda.synchronize()
        if r > 0: best = min(best, e0.elapsed_time(e1))
    waves = -(-B // 296)
    print("B=%5d n=%4d M=%d: %8.3f ms  -> %7.2f us per matrix-slot (%d waves)" % (B, n, M, best, best * 1e3 / waves, waves), flush=True)
for n in (64, 128, 192, 256, 512, 1000):
    timeit(1, n, 0)
    timeit(296 * 8, n, 0)
