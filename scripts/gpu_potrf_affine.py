"""GPU-box probe: cost and correctness of the fused affine path of the Cholesky kernel."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import starry_process_b200 as spb
from starry_process_b200 import _lib
dev = torch.device("cuda:0")
c = spb.get_context(0); lib, ctx = c.lib, c.handle
P = lambda x: ctypes.c_void_p(x.data_ptr()) if x is not None else None
B, n = 1184, 1000
A = torch.randn(B, n, 32, dtype=torch.float64, device=dev)
K0 = torch.bmm(A, A.transpose(1, 2)) / 32 + torch.eye(n, dtype=torch.float64, device=dev)
R0 = torch.randn(B, 1, n, dtype=torch.float64, device=dev)
q = torch.rand(B, n, dtype=torch.float64, device=dev) * 1e-3
scal = torch.zeros(B, 4, dtype=torch.float64, device=dev); scal[:, 0] = 1.01; scal[:, 1] = 2e-3; scal[:, 2] = 1e-3
dg = torch.full((1,), 1e-6, dtype=torch.float64, device=dev)
off = torch.full((1,), 1e-5, dtype=torch.float64, device=dev)
def run(mode):
    ll = torch.zeros(B, dtype=torch.float64, device=dev); info = torch.zeros(B, dtype=torch.int32, device=dev)
    best = 1e9
    for r in range(4):
        K = K0.clone(); R = R0.clone(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode == "plain":
            _lib.check(lib.spb_cholesky_lnlike(ctx, B, n, P(K), n, n * n, 1, P(R), n, n, P(ll), None, None, P(info), None))
        else:
            af = _lib.Affine()
            if mode in ("full", "norm"):
                af.scal, af.q = P(scal), P(q)
            if mode in ("full", "noise"):
                af.diag, af.diag_kind, af.diag_stride = P(dg), 0, 0
                af.offset, af.offset_stride = P(off), 0
            _lib.check(lib.spb_cholesky_lnlike_affine(ctx, B, n, P(K), n, n * n, ctypes.byref(af), 1, P(R), n, n, P(ll), None, None, P(info), None))
        e1.record(); torch.cuda.synchronize()
        if r > 0: best = min(best, e0.elapsed_time(e1))
    print("%-6s %.3f ms  ll[0]=%.6f" % (mode, best, float(ll[0])), flush=True)
    return ll
l0 = run("plain"); l1 = run("ident"); run("noise"); run("norm"); l2 = run("full")
print("identity affine == plain:", float((l0 - l1).abs().max()))
# reference for 'full' on 2 matrices
Kf = scal[:2, 0, None, None] * K0[:2] + scal[:2, 1, None, None] * ((1 - q[:2, :, None]) * (1 - q[:2, None, :])) - scal[:2, 2, None, None] * (q[:2, :, None] * q[:2, None, :]) + off + torch.diag_embed(dg.expand(2, n))
L = torch.linalg.cholesky(Kf); y = torch.linalg.solve_triangular(L, R0[:2].transpose(1, 2), upper=False)
ref = -0.5 * (y ** 2).sum((1, 2)) - torch.log(torch.diagonal(L, dim1=1, dim2=2)).sum(1) - 0.5 * n * np.log(2 * np.pi)
print("full affine vs torch:", (l2[:2] - ref).abs().cpu().numpy(), ref.cpu().numpy())
