"""GPU-box sanity script (development aid): potrf/lnlike kernel vs NumPy, DMMA and cuBLAS FP64 peaks."""
import ctypes, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from starry_process_b200 import _lib

lib = _lib.load(allow_missing_symbols=True)
ctx = ctypes.c_void_p()
_lib.check(lib.spb_create(0, None, 0, ctypes.byref(ctx)))
dev = torch.device("cuda:0")
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None


def run(B, n, M, seed=0, ld=None, check=True):
    rng = np.random.default_rng(seed)
    ld = ld or (n + (n % 2))
    A = rng.standard_normal((B, n, 40))
    Kh = A @ A.transpose(0, 2, 1) / 40 + 0.5 * np.eye(n)[None]
    Kp = np.zeros((B, n, ld)); Kp[:, :, :n] = Kh
    iu = np.triu_indices(n, 1); Kp[:, iu[0], iu[1]] = np.nan  # the kernel must never read above the diagonal
    K = torch.tensor(Kp, device=dev)
    R = None; Rh = None
    if M > 0:
        Rh = rng.standard_normal((B, M, n))
        Rp = np.zeros((B, M, ld)); Rp[:, :, :n] = Rh
        R = torch.tensor(Rp, device=dev)
    ll = torch.zeros(B, dtype=torch.float64, device=dev)
    quad = torch.zeros(B, max(M, 1), dtype=torch.float64, device=dev)
    logdet = torch.zeros(B, dtype=torch.float64, device=dev)
    info = torch.zeros(B, dtype=torch.int32, device=dev)
    st = lib.spb_cholesky_lnlike(ctx, B, n, P(K), ld, n * ld, M, P(R), ld, M * ld, P(ll), P(quad), P(logdet), P(info), None)
    _lib.check(st)
    torch.cuda.synchronize()
    if not check:
        return
    Lg = K.cpu().numpy()[:, :, :n]
    err = 0; errq = 0; errl = 0
    for b in range(B):
        L = np.linalg.cholesky(Kh[b])
        err = max(err, np.abs(np.tril(Lg[b]) - L).max())
        ld_ref = np.log(np.diag(L)).sum()
        errl = max(errl, abs(ld_ref - logdet[b].item()))
        if M > 0:
            import scipy.linalg as sl
            y = sl.solve_triangular(L, Rh[b].T, lower=True)
            q_ref = (y ** 2).sum(0)
            errq = max(errq, np.abs(q_ref - quad[b].cpu().numpy()[:M]).max() / np.abs(q_ref).max())
            yg = R.cpu().numpy()[b][:, :n]
            errq = max(errq, np.abs(yg - y.T).max())
            ll_ref = -0.5 * q_ref.sum() - M * ld_ref - 0.5 * n * M * np.log(2 * np.pi)
            errl = max(errl, abs(ll_ref - ll[b].item()) / abs(ll_ref))
    print("B=%d n=%d M=%d ld=%d: |L-Lref|=%.2e quad/y err=%.2e logdet/lnlike err=%.2e info=%s" % (B, n, M, ld, err, errq, errl, info.cpu().numpy().tolist()[:4]), flush=True)


for (B, n, M) in [(2, 64, 0), (2, 64, 1), (3, 100, 2), (2, 128, 1), (2, 130, 5), (2, 200, 1), (2, 257, 140), (3, 1000, 1), (1, 1000, 8), (1, 1, 1), (2, 7, 2), (1, 1026, 3)]:
    run(B, n, M)

# non-PD handling
K = torch.eye(100, dtype=torch.float64, device=dev).repeat(2, 1, 1).contiguous(); K[1, 50, 50] = -1.0
R = torch.ones(2, 1, 100, dtype=torch.float64, device=dev)
ll = torch.zeros(2, dtype=torch.float64, device=dev); info = torch.zeros(2, dtype=torch.int32, device=dev)
_lib.check(lib.spb_cholesky_lnlike(ctx, 2, 100, P(K), 100, 100 * 100, 1, P(R), 100, 100, P(ll), None, None, P(info), None))
torch.cuda.synchronize(); print("nonPD:", ll.cpu().numpy(), info.cpu().numpy())

# solve_rows
rng = np.random.default_rng(5); n = 1000; M = 300
A = rng.standard_normal((n, 50)); Kh = A @ A.T / 50 + 0.5 * np.eye(n); L = np.linalg.cholesky(Kh)
Lt = torch.tensor(L, device=dev); Rh = rng.standard_normal((M, n)); R = torch.tensor(Rh, device=dev)
quad = torch.zeros(M, dtype=torch.float64, device=dev)
_lib.check(lib.spb_cholesky_solve_rows(ctx, n, P(Lt), n, M, P(R), n, P(quad), None)); torch.cuda.synchronize()
import scipy.linalg as sl
y = sl.solve_triangular(L, Rh.T, lower=True)
print("solve_rows: y err %.2e quad err %.2e" % (np.abs(R.cpu().numpy() - y.T).max(), np.abs(quad.cpu().numpy() - (y ** 2).sum(0)).max()))

# timing
def timeit(B, n, M, reps=3):
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
    fl = B * (n ** 3 / 3 + n * n * M)
    print("time B=%d n=%d M=%d: %.3f ms -> %.2f TFLOP/s (%.0f matrices/s)" % (B, n, M, best, fl / best / 1e9, B / best * 1e3), flush=True)

for (B, n, M) in [(296, 1000, 1), (1184, 1000, 1), (2368, 1000, 1), (296, 1024, 1), (148, 4096, 1), (1, 1000, 0)]:
    timeit(B, n, M)

tf = ctypes.c_double(); ms = ctypes.c_double()
for it in (20000, 100000):
    _lib.check(lib.spb_dmma_peak(ctx, it, ctypes.byref(tf), ctypes.byref(ms)))
    print("DMMA peak microbench: iters=%d %.2f TFLOP/s (%.2f ms)" % (it, tf.value, ms.value), flush=True)
# cuBLAS fp64 GEMM
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b); torch.cuda.synchronize(); best = 1e9
    for r in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = torch.matmul(a, b); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print("cuBLAS fp64 GEMM n=%d: %.2f TFLOP/s" % (n, 2 * n ** 3 / best / 1e9), flush=True)
# torch.linalg.cholesky batched fp64 for context
A = torch.randn(296, 1000, 32, dtype=torch.float64, device=dev); K0 = torch.bmm(A, A.transpose(1, 2)) / 32 + torch.eye(1000, dtype=torch.float64, device=dev)
torch.linalg.cholesky(K0); torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); torch.linalg.cholesky(K0); e1.record(); torch.cuda.synchronize()
print("torch.linalg.cholesky (cuSOLVER/MAGMA) 296x1000: %.2f ms" % e0.elapsed_time(e1))
