"""GPU-box timing of the design-matrix phase (BASELINE configs[4]): nt = 1e5 timestamps x I
inclinations, plus the standalone tensordotRz and sample_ylm bandwidth kernels.  CUDA events on the
launching stream, best of `reps` after warm-up."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import starry_process_b200 as spb
from starry_process_b200 import _lib

dev = torch.device("cuda:0")
ctx = spb.get_context(0)
lib, h = ctx.lib, ctx.handle
P = lambda x: ctypes.c_void_p(x.data_ptr())
gp = spb.StarryProcess(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0)
rta1 = gp._rTA1([0.4, 0.26])

def time_design(I, nt, reps=5):
    t = torch.linspace(0, 40, nt, dtype=torch.float64, device=dev)
    inc = torch.arccos(torch.rand(I, dtype=torch.float64, device=dev))
    per = torch.ones(I, dtype=torch.float64, device=dev)
    A = torch.empty(I, nt, 256, dtype=torch.float64, device=dev)
    nb = lib.spb_design_matrix_workspace_bytes(h, I, nt)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    best = 1e9
    for r in range(reps + 2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.spb_design_matrix(h, I, nt, P(t), P(inc), P(per), P(rta1), 0, P(A), P(ws), nb, None))
        e1.record(); torch.cuda.synchronize()
        if r >= 2: best = min(best, e0.elapsed_time(e1))
    gb = I * nt * 2056.0 / 1e9
    print("design matrix I=%d nt=%d: %.3f ms  %.1f GB/s algorithmic (%.2f GB), %.2f Mrows/s" % (
        I, nt, best, gb / best * 1e3, gb, I * nt / best / 1e3), flush=True)
    return A

A = time_design(64, 100000)
del A
time_design(16, 100000)
time_design(4096, 1000)
time_design(1, 1000)
# tensordotRz standalone
K = 2000000
M = torch.randn(K, 256, dtype=torch.float64, device=dev); th = torch.rand(K, dtype=torch.float64, device=dev) * 6.28
f = torch.empty_like(M); best = 1e9
for r in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); _lib.check(lib.spb_tensordotRz(h, K, P(M), P(th), P(f), None)); e1.record(); torch.cuda.synchronize()
    if r >= 1: best = min(best, e0.elapsed_time(e1))
print("tensordotRz K=%d: %.3f ms  %.1f GB/s (4104 B/row)" % (K, best, K * 4104.0 / best / 1e6), flush=True)
del M, f
# sample_ylm
ns = 1000000
L = gp.cho_cov_ylm
u = torch.randn(1, ns, 256, dtype=torch.float64, device=dev); y = torch.empty_like(u); best = 1e9
for r in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); _lib.check(lib.spb_sample_ylm(h, 1, ns, P(gp._mean_ylm), P(L.contiguous()), P(u), P(y), None)); e1.record(); torch.cuda.synchronize()
    if r >= 1: best = min(best, e0.elapsed_time(e1))
print("sample_ylm ns=%d: %.3f ms  %.1f GB/s (4096 B/draw read+write), %.2f TFLOP/s" % (ns, best, ns * 4096.0 / best / 1e6, ns * 2 * 65536 / best / 1e9), flush=True)
