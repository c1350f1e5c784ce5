"""GPU-box probe: phase timers of potrf_lnlike_kernel (debug build libspb200_prof.so, -DSPB_POTRF_PROF).
Prints, for warp 0 and warp 3 (both diagonal-block warps; warp 3 owns the last two sub-panels), the share of SM cycles per phase."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["SPB200_LIB"] = os.path.join(ROOT, "starry_process_b200", "libspb200_prof.so")
sys.path.insert(0, ROOT)
import numpy as np, torch
import starry_process_b200 as spb
from starry_process_b200 import _lib
dev = torch.device("cuda:0")
c = spb.get_context(0); lib, ctx = c.lib, c.handle
lib.spb_potrf_prof.restype = ctypes.c_int
lib.spb_potrf_prof.argtypes = [ctypes.c_void_p]
P = lambda x: ctypes.c_void_p(x.data_ptr())
NAMES = ["issue01", "init_acc", "wait_chunk0", "kloop", "sync+writeback", "trsm", "store", "panel_sync",
         "reduce", "pre_issue", "matrices", "d:tile8", "d:defer+bar1", "d:trsm+bar2", "d:crit_upd", "next_issue"]
KEYS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 12, 13, 14]
def run(B, n, M=1):
    A = torch.randn(B, n, 32, dtype=torch.float64, device=dev)
    K0 = torch.bmm(A, A.transpose(1, 2)) / 32 + torch.eye(n, dtype=torch.float64, device=dev)
    R0 = torch.randn(B, max(M, 1), n, dtype=torch.float64, device=dev)
    ll = torch.zeros(B, dtype=torch.float64, device=dev); info = torch.zeros(B, dtype=torch.int32, device=dev)
    out = (ctypes.c_ulonglong * 36)()
    for r in range(2):
        K = K0.clone(); R = R0.clone(); torch.cuda.synchronize()
        lib.spb_potrf_prof(out)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.spb_cholesky_lnlike(ctx, B, n, P(K), n, n * n, M, P(R), n, max(M, 1) * n, P(ll), None, None, P(info), None))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        lib.spb_potrf_prof(out)
    v = np.array(list(out)[:32], dtype=np.float64).reshape(2, 16)
    ck = list(out)[32:]
    print("B=%d n=%d: %.3f ms; CTA 0 ran %.3f ms at an effective SM clock of %.0f MHz" % (
        B, n, ms, (ck[3] - ck[1]) * 1e-6, (ck[2] - ck[0]) / max(ck[3] - ck[1], 1) * 1e3))
    for w in range(2):
        nm = v[w, 10]; tot = v[w, KEYS].sum()
        print("  warp %d: %.0f matrices, %.1f kclk per matrix" % (3 * w, nm, tot / nm / 1e3))
        for k in KEYS:
            print("     %-12s %9.1f kclk/matrix  %5.1f%%" % (NAMES[k], v[w, k] / nm / 1e3, 100 * v[w, k] / tot))
    sys.stdout.flush()
run(1, 1000)
run(148, 1000)
run(296 * 4, 1000)
run(296 * 12, 1000)
