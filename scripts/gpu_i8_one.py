"""One launch of spb_cholesky_lnlike_i8 (default B = 148, 8 planes, nt = 1000, M = 1) for ncu:
   python scripts/gpu_i8_one.py [B [planes [nt]]]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import starry_process_b200 as spb
from starry_process_b200 import _lib
dev = torch.device("cuda:0")
c = spb.get_context(0); lib, ctx = c.lib, c.handle
P = lambda x: ctypes.c_void_p(x.data_ptr())
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
planes = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
M = 1
A = torch.randn(B, n, 32, dtype=torch.float64, device=dev)
K = torch.bmm(A, A.transpose(1, 2)) / 32
R = 0.01 * torch.randn(B, 1, n, dtype=torch.float64, device=dev)
dg = torch.full((1,), 1e-4, dtype=torch.float64, device=dev)
af = _lib.Affine(); af.diag, af.diag_kind, af.diag_stride = dg.data_ptr(), 0, 0
ll = torch.zeros(B, dtype=torch.float64, device=dev); info = torch.zeros(B, dtype=torch.int32, device=dev)
nb = lib.spb_cholesky_i8_workspace_bytes(B, n, M, planes)
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
for rep in range(2):
    _lib.check(lib.spb_cholesky_lnlike_i8(ctx, B, n, P(K), n, n * n, ctypes.byref(af), M, P(R.clone()), n, n,
                                          P(ll), None, None, P(info), planes, 0.0, P(ws), nb, None))
torch.cuda.synchronize()
print(ll[:3].tolist(), info[:3].tolist())
