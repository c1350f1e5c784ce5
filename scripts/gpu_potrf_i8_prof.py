"""GPU-box probe: phase timers of potrf_i8_kernel (debug build libspb200_prof.so, -DSPB_POTRF_PROF)."""
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
CN = {0: "init_acc", 1: "wait tmem_full", 2: "epilogue", 3: "potf2+bar (diag)", 4: "trsm", 5: "stores+slicing",
      6: "fence+barrier", 7: "reduce", 8: "set-up (scales)"}
IN = {0: "issuer: wait tmem_empty", 1: "issuer: wait full", 2: "issuer: issue", 8: "producer: wait stored",
      9: "producer: wait empty", 10: "producer: issue"}
def run(B, n, M=1, planes=8):
    A = torch.randn(B, n, 32, dtype=torch.float64, device=dev)
    K0 = torch.bmm(A, A.transpose(1, 2)) / 32
    ld = n + (n & 1)
    K = torch.zeros(B, n, ld, dtype=torch.float64, device=dev); K[:, :, :n] = K0
    R = torch.zeros(B, max(M, 1), ld, dtype=torch.float64, device=dev)
    R[:, :, :n] = 0.01 * torch.randn(B, max(M, 1), n, dtype=torch.float64, device=dev)
    dg = torch.full((1,), 1e-4, dtype=torch.float64, device=dev)
    af = _lib.Affine(); af.diag, af.diag_kind, af.diag_stride = dg.data_ptr(), 0, 0
    ll = torch.zeros(B, dtype=torch.float64, device=dev); info = torch.zeros(B, dtype=torch.int32, device=dev)
    nb = lib.spb_cholesky_i8_workspace_bytes(B, n, M, planes)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    out = (ctypes.c_ulonglong * 36)()
    for r in range(2):
        Rc = R.clone(); torch.cuda.synchronize()
        lib.spb_potrf_prof(out)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.spb_cholesky_lnlike_i8(ctx, B, n, P(K), ld, n * ld, ctypes.byref(af), M, P(Rc), ld, max(M, 1) * ld,
                                              P(ll), None, None, P(info), planes, 0.0, P(ws), nb, None))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        lib.spb_potrf_prof(out)
    v = np.array(list(out)[:32], dtype=np.float64).reshape(2, 16)
    nm = v[0, 10]
    print("B=%d n=%d planes=%d: %.3f ms, %.0f matrices" % (B, n, planes, ms, nm))
    tot = sum(v[0, k] for k in CN)
    print("  compute thread 0: %.1f kclk per matrix" % (tot / nm / 1e3))
    for k, name in CN.items():
        print("     %-26s %9.1f kclk/matrix  %5.1f%%" % (name, v[0, k] / nm / 1e3, 100 * v[0, k] / tot))
    for k, name in {11: "d:tile8", 12: "d:defer+bar1", 13: "d:trsm+bar2", 14: "d:crit_upd",
                    15: "d:tile8 of panel 0 only (no MMA / TMA in flight)",
                    9: "init_acc of panel 0 only (8 of the 68 tiles)"}.items():
        print("       %-24s %9.1f kclk/matrix" % (name, v[0, k] / nm / 1e3))
    for k, name in IN.items():
        print("     %-26s %9.1f kclk/matrix" % (name, v[1, k] / nm / 1e3))
    sys.stdout.flush()
#run(148, 1000)
run(1184, 1000, planes=78)
#run(1184, 1000, planes=7)
