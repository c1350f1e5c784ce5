"""ncu driver: potrf_lnlike_kernel on tiny matrices (n = 64: the diagonal-block path only)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import starry_process_b200 as spb
from starry_process_b200 import _lib
dev = torch.device("cuda:0")
c = spb.get_context(0); lib, ctx = c.lib, c.handle
P = lambda x: ctypes.c_void_p(x.data_ptr())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = 2368
A = torch.randn(B, n, 32, dtype=torch.float64, device=dev)
K = torch.bmm(A, A.transpose(1, 2)) / 32 + torch.eye(n, dtype=torch.float64, device=dev)
ll = torch.zeros(B, dtype=torch.float64, device=dev); info = torch.zeros(B, dtype=torch.int32, device=dev)
_lib.check(lib.spb_cholesky_lnlike(ctx, B, n, P(K), n, n * n, 0, None, n, 0, P(ll), None, None, P(info), None))
torch.cuda.synchronize()
print("done", float(ll[0]))
