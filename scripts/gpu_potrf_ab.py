"""GPU-box probe: the two tile geometries of the batched Cholesky (cholesky_tile = 128: 8 warps x 2 CTAs
per SM; 64: 4 warps x 3 CTAs per SM, packed L_jj) -- correctness against LAPACK, bit-equality of the
factors between the two, and matrices/s over batch sizes."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import starry_process_b200 as spb
from starry_process_b200 import _lib
dev = torch.device("cuda:0")
c = spb.get_context(0); lib, ctx = c.lib, c.handle
P = lambda x: ctypes.c_void_p(x.data_ptr())

def run(K0, R0, n, M, tile, reps=1, prefetch=1):
    c.set_option("cholesky_tile", tile)
    c.set_option("cholesky_prefetch", prefetch)
    B = K0.shape[0]
    ld = K0.shape[2]
    best = 1e30
    for r in range(reps):
        K, R = K0.clone(), R0.clone()
        ll = torch.zeros(B, dtype=torch.float64, device=dev); info = torch.zeros(B, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.spb_cholesky_lnlike(ctx, B, n, P(K), ld, n * ld, M, P(R) if M else None, ld, max(M, 1) * ld, P(ll), None, None, P(info), None))
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return K, R, ll, info, best

rng = np.random.default_rng(0)
print("== correctness")
for (B, n, M) in [(2, 1, 1), (2, 7, 2), (3, 64, 0), (2, 100, 3), (2, 129, 1), (2, 257, 140), (5, 1000, 1), (2, 1000, 70), (1, 2049, 1)]:
    ld = n + (n & 1)
    A = rng.standard_normal((B, n, 40))
    Kh = A @ A.transpose(0, 2, 1) / 40 + 0.5 * np.eye(n)[None]
    Kp = np.zeros((B, n, ld)); Kp[:, :, :n] = Kh
    Rh = rng.standard_normal((B, max(M, 1), n)); Rp = np.zeros((B, max(M, 1), ld)); Rp[:, :, :n] = Rh
    K0, R0 = torch.tensor(Kp, device=dev), torch.tensor(Rp, device=dev)
    c.set_option("cholesky_cluster", 0)
    out = {}
    for tile in (128, 64, 163, 164, 165):
        for pf in (0, 1):
            K, R, ll, info, _ = run(K0, R0, n, M, tile, prefetch=pf)
            out[(tile, pf)] = (torch.tril(K[:, :, :n]), R.clone(), ll.clone())
    for tile in (128, 64):
        errL = max(np.abs(out[(tile, 1)][0][b].cpu().numpy() - np.linalg.cholesky(Kh[b])).max() for b in range(B))
        print("  B=%d n=%d M=%d tile=%d: max |L - LAPACK| = %.2e info=%s" % (B, n, M, tile, errL, info.cpu().numpy().tolist()))
    ref = out[(128, 0)]
    print("     all variants bitwise equal (factor, rhs rows):",
          all(torch.equal(v[0], ref[0]) and torch.equal(v[1], ref[1]) for v in out.values()),
          " max |dlnlike|/|lnlike| = %.2e" % max(float(((v[2] - ref[2]).abs() / ref[2].abs().clamp_min(1e-300)).max()) for v in out.values()))
    c.set_option("cholesky_cluster", 1)

print("== throughput, nt = 1000, M = 1")
n, M = 1000, 1
Bmax = 4096
gen = torch.Generator(device=dev).manual_seed(1)
A = torch.randn(Bmax, n, 32, dtype=torch.float64, device=dev, generator=gen)
Kall = torch.bmm(A, A.transpose(1, 2)) / 32 + torch.eye(n, dtype=torch.float64, device=dev)
del A
Rall = torch.randn(Bmax, 1, n, dtype=torch.float64, device=dev, generator=gen)
flop = n ** 3 / 3.0 + n ** 2
c.set_option("cholesky_cluster", 0)
for B in (148, 296, 444, 512, 592, 888, 1184, 2368, 4096):
    line = "  B=%5d" % B
    for tile in (128, 64, 163, 164, 165):
        _, _, _, _, ms = run(Kall[:B], Rall[:B], n, M, tile, reps=3, prefetch=1)
        line += "  t%d %7.3f ms %5.2f TF" % (tile, ms, B * flop / ms / 1e9)
    print(line, flush=True)
c.set_option("cholesky_cluster", 1)
del Kall, Rall
print("== throughput, nt = 4096, M = 1")
n = 4096
B = 148
gen = torch.Generator(device=dev).manual_seed(2)
A = torch.randn(B, n, 32, dtype=torch.float64, device=dev, generator=gen)
Kall = torch.bmm(A, A.transpose(1, 2)) / 32 + torch.eye(n, dtype=torch.float64, device=dev)
Rall = torch.randn(B, 1, n, dtype=torch.float64, device=dev, generator=gen)
flop = n ** 3 / 3.0 + n ** 2
for tile in (128, 64, 163, 164, 165):
    _, _, _, _, ms = run(Kall, Rall, n, 1, tile, reps=2, prefetch=1)
    print("  B=%d tile %4d: %8.2f ms  %5.2f TF/s" % (B, tile, ms, B * flop / ms / 1e9))
c.set_option("cholesky_tile", 64)
