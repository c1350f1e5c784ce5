"""GPU-box probe: batched Cholesky throughput for small / mid-sized batches (cluster sizes 8, 4, 2
against the one-CTA-per-matrix kernel; SPB_NO_CLUSTER=1 selects the latter)."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1:
    for flag in ("0", "1"):
        env = dict(os.environ)
        if flag == "1":
            env["SPB_NO_CLUSTER"] = "1"
        subprocess.run([sys.executable, __file__, flag], env=env)
    sys.exit(0)
sys.path.insert(0, ROOT)
import numpy as np, torch
import starry_process_b200 as spb
from starry_process_b200 import _lib
dev = torch.device("cuda:0")
c = spb.get_context(0); lib, ctx = c.lib, c.handle
P = lambda x: ctypes.c_void_p(x.data_ptr())
tag = "one CTA per matrix" if sys.argv[1] == "1" else "clusters"
ref = {}
for B, n in ((1, 1000), (4, 1000), (18, 1000), (30, 1000), (37, 1000), (60, 1000), (74, 1000), (2, 4096)):
    torch.manual_seed(B)
    A = torch.randn(B, n, 48, dtype=torch.float64, device=dev)
    K0 = torch.bmm(A, A.transpose(1, 2)) / 48 + torch.eye(n, dtype=torch.float64, device=dev)
    R0 = torch.randn(B, 1, n, dtype=torch.float64, device=dev)
    ll = torch.zeros(B, dtype=torch.float64, device=dev); info = torch.zeros(B, dtype=torch.int32, device=dev)
    best = 1e9
    for r in range(4):
        K = K0.clone(); R = R0.clone(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.spb_cholesky_lnlike(ctx, B, n, P(K), n, n * n, 1, P(R), n, n, P(ll), None, None, P(info), None))
        e1.record(); torch.cuda.synchronize()
        if r: best = min(best, e0.elapsed_time(e1))
    Lr = torch.linalg.cholesky(K0)
    err = float((torch.tril(K) - Lr).abs().max())
    print("%-18s B=%3d n=%4d: %7.3f ms  (%6.0f matrices/s)  |L-Lref|=%.1e  ll[0]=%.9f" % (
        tag, B, n, best, B / best * 1e3, err, ll[0].item()), flush=True)
