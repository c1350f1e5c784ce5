"""Accuracy study (test infrastructure, CPU only): left-looking blocked Cholesky of the bench
covariances with the trailing update  P = K - L L^T  evaluated the way an INT8 tensor-core
emulation of FP64 would (Ozaki-style): every row of L is scaled by a power of two taken from the
diagonal of K (|L_ik| <= sqrt(K_ii)), cut into signed 7-bit digits, the digit planes are multiplied
exactly (integer), plane pairs (s, t) with s + t <= D are kept, and the groups are combined in
fp64.  Prints the lnlike difference against LAPACK for a number of digit planes."""
import os
import sys
import numpy as np
import scipy.linalg as sl

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sp_oracle  # noqa: E402
import bench  # noqa: E402

NB = 64


def digits(X, e_row, S, bits=7):
    """Signed-digit planes of X / 2^e_row (|.| <= 1): X ~= 2^e * sum_s d_s 2^{-bits (s+1)}."""
    x = np.ldexp(X, -e_row[:, None])          # exact scaling
    planes = []
    for s in range(S):
        x = np.ldexp(x, bits)
        d = np.rint(x)
        planes.append(d)
        x = x - d                              # exact (Sterbenz-like: |x - d| <= 1/2)
    return planes                              # each within [-2^bits, 2^bits] (first plane), then [-64, 64]


def sliced_product(A_planes, B_planes, D, bits=7):
    """sum over plane pairs s + t <= D of A_s B_t^T 2^{-bits (s + t + 2)}, groups summed small -> large."""
    acc = None
    for d in range(D, -1, -1):
        G = None
        for s in range(d + 1):
            t = d - s
            if s >= len(A_planes) or t >= len(B_planes):
                continue
            P = A_planes[s] @ B_planes[t].T      # exact: small integers
            G = P if G is None else G + P
        term = np.ldexp(G, -bits * (d + 2))
        acc = term if acc is None else acc + term
    return acc


def blocked_cholesky(K, mode, S=8, D=None, bits=7):
    n = K.shape[0]
    L = np.zeros_like(K)
    if D is None:
        D = S - 1
    # per-row exponent: |L_ik| <= sqrt(K_ii) < 2^e
    e_row = (np.floor(np.log2(np.sqrt(np.diag(K)))) + 1).astype(np.int64)
    if bits == 8:
        e_row = e_row + 1      # radix 256: every digit uses the full int8 range, so |x| <= 1/2
    planes = [np.zeros((n, 0)) for _ in range(S)]
    for c0 in range(0, n, NB):
        c1 = min(n, c0 + NB)
        P = K[c0:, c0:c1].copy()
        if c0 > 0:
            if mode == "fp64":
                P -= L[c0:, :c0] @ L[c0:c1, :c0].T
            else:
                A = [p[c0:, :] for p in planes]
                Bp = [p[c0:c1, :] for p in planes]
                acc = sliced_product(A, Bp, D, bits)
                P -= np.ldexp(acc, (e_row[c0:, None] + e_row[None, c0:c1]))
        Ljj = np.linalg.cholesky(P[: c1 - c0])
        L[c0:c1, c0:c1] = Ljj
        if c1 < n:
            L[c1:, c0:c1] = sl.solve_triangular(Ljj, P[c1 - c0:].T, lower=True).T
        if mode != "fp64":
            newp = digits(L[:, c0:c1], e_row, S, bits)
            planes = [np.concatenate([p, q], axis=1) for p, q in zip(planes, newp)]
    return L


def lnlike_from(L, r):
    w = sl.solve_triangular(L, r, lower=True)
    return -0.5 * w @ w - np.sum(np.log(np.diag(L))) - 0.5 * len(r) * np.log(2 * np.pi)


def main():
    nd = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    for prior in ("narrow", "full"):
        hp, t, flux, _ = bench.synthetic_inputs(64, 1234 if prior == "narrow" else 4321, prior)
        for b in range(nd):
            gp = sp_oracle.OracleProcess(r=hp["r"][b], c=hp["c"][b], n=hp["n"][b], mu=hp["mu"][b],
                                         sigma=hp["sigma"][b], normalized=True,
                                         marginalize_over_inclination=True)
            ll, K, Lref = gp.log_likelihood(t, flux, 1e-6, u=(0.4, 0.26), return_parts=True)
            if not np.isfinite(ll):
                continue
            r = flux - gp.mean(t, u=(0.4, 0.26))
            ref = lnlike_from(np.linalg.cholesky(K), r)
            Kq = K.astype(np.longdouble)
            out = ["%s #%d cond %.1e" % (prior, b, np.linalg.cond(K))]
            out.append("blocked fp64 %.1e" % abs((lnlike_from(blocked_cholesky(K, "fp64"), r) - ref) / ref))
            for S in (7, 8):
                v = lnlike_from(blocked_cholesky(K, "int8", S=S), r)
                out.append("S=%d %.1e" % (S, abs((v - ref) / ref)))
            v = lnlike_from(blocked_cholesky(K, "int8", S=7, bits=8), r)
            out.append("S=7 radix 256 %.1e" % abs((v - ref) / ref))
            print("  ".join(out), flush=True)


if __name__ == "__main__":
    main()
