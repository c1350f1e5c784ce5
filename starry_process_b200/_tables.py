"""Host-side constant tables for libspb200 (hyperparameter-independent, computed once per process).

This is the product's mirror of the NumPy precomputations that the reference performs at
graph-build time, re-organised for the GPU kernels ("constant tables"):

  * Spot profile operator ``Bp``                       -- size.py:10-43
  * polynomial-coefficient Wigner tensors ``R[l]``     -- wigner.py:192-372
  * latitude: range basis ``Z`` of the rank-31 moment matrix and the folded tensors
    ``H = R_lat[:, m=0, :] . Z``, ``R0 = R_lat[:, m=0, :]``  -- latitude.py:203-211, integrals.py:116-124
  * longitude moments / tensors ``t_lon, T_lon``       -- longitude.py:9-49, integrals.py:116-124,
    math.py:121-139 (eigh + 1e-15 clip of the constant matrix Q_lon)
  * flux: inclination-marginalisation integrals ``wnp, Wnp`` (flux.py:107-179) folded with
    ``Rx(pi/2)`` into the 16 quadratic forms ``Omega_m`` used by the marginal-kernel GEMM
  * Gauss-Legendre nodes for the limb-darkened flux operator.

``wigner_poly`` and ``rx_numeric`` are a PORT of the reference's table construction (the same
recurrences in the same order, wigner.py:192-372 / ops/include/wigner.h:37-284, so that the constant
tensors come out bit-compatible with the ones the reference builds); every other table is this
package's own re-organisation.  They run once per process on the host and are not on the hot path.

Only NumPy/SciPy; nothing here touches the GPU.  The packed blob layout is shared with
``csrc/spb_tables.h``.
"""
import math
import os

import numpy as np
from scipy.special import gamma as _gamma
from scipy.special import hyp2f1 as _hyp2f1
from scipy.special import legendre as _legendre

YDEG = 15
N = (YDEG + 1) ** 2
NEIG = 2 * YDEG + 1
NWIG = ((YDEG + 1) * (2 * YDEG + 1) * (2 * YDEG + 3)) // 3
SPTS = 1000
HERE = os.path.dirname(os.path.abspath(__file__))

# (name, number of doubles); order defines the blob layout -- keep in sync with csrc/spb_tables.h
LAYOUT = [
    ("THETA", SPTS),
    ("BP", (YDEG + 1) * SPTS),
    ("LAT_Z", N * 32),
    ("LAT_H", N * 32),
    ("LAT_R0", NWIG),
    ("LAT_FAC", 46376),
    ("LAT_FACOFF", 31 * 31 + 1),
    ("LON_T1", NWIG),
    ("LON_T", NEIG * NWIG),
    ("FLUX_WNP", NWIG),
    ("FLUX_W", N * N),
    ("FLUX_OMEGA", 31 * N * N),
    ("RX90", NWIG),
    ("RX90_NZ", 1372),
    ("GL_X", 32),
    ("GL_W", 32),
    ("LAMBDA", N),
]


def offsets():
    off = {}
    pos = 0
    for name, cnt in LAYOUT:
        off[name] = pos
        pos += cnt
    off["_TOTAL"] = pos
    return off


def nwig(l):
    return ((l + 1) * (2 * l + 1) * (2 * l + 3)) // 3


def _lm():
    l = np.concatenate([np.full(2 * ll + 1, ll) for ll in range(YDEG + 1)])
    m = np.concatenate([np.arange(-ll, ll + 1) for ll in range(YDEG + 1)])
    return l, m


# ----------------------------------------------------------------------------------------------
# polynomial Wigner tensors (Alvarez Collado et al. recurrences on coefficient vectors).  The
# operation ORDER follows wigner.py:192-292 exactly: the coefficients reach 7.8e7 at l = 15 and
# cancel heavily downstream, so a different (equally valid) rounding sequence would move
# cov_ylm at the 1e-6 relative level ("numerical fragility").
# ----------------------------------------------------------------------------------------------
def _pmul(x1, x2):
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    out = np.zeros(x1.size + x2.size - 1)
    for i in range(x1.size):          # accumulate in ascending order of the first factor's index
        out[i:i + x2.size] += x1[i] * x2
    return out


def wigner_poly(ydeg, c1, s1, c3, s3):
    """R[l][m', m, k]: coefficient of sin(phi/2)^(2l-k) cos(phi/2)^k (wigner.py:295-372)."""
    rt2 = math.sqrt(2.0)
    D = [np.full((2 * l + 1,) * 3, np.nan) for l in range(ydeg + 1)]
    R = [np.full((2 * l + 1,) * 3, np.nan) for l in range(ydeg + 1)]
    D[0][0, 0] = [1]
    R[0][0, 0] = [1]
    D[1][2, 2] = [0, 0, 1]
    D[1][2, 1] = [0, -rt2, 0]
    D[1][2, 0] = [1, 0, 0]
    D[1][1, 2] = -D[1][2, 1]
    D[1][1, 1] = D[1][2, 2] - D[1][2, 0]
    D[1][1, 0] = D[1][2, 1]
    D[1][0, 2] = D[1][2, 0]
    D[1][0, 1] = D[1][1, 2]
    D[1][0, 0] = D[1][2, 2]
    cosag = c1 * c3 - s1 * s3
    cosamg = c1 * c3 + s1 * s3
    sinag = s1 * c3 + c1 * s3
    sinamg = s1 * c3 - c1 * s3
    R[1][1, 1] = D[1][1, 1]
    R[1][2, 1] = rt2 * D[1][1, 2] * c1
    R[1][0, 1] = rt2 * D[1][1, 2] * s1
    R[1][1, 2] = rt2 * D[1][2, 1] * c3
    R[1][1, 0] = -rt2 * D[1][2, 1] * s3
    R[1][2, 2] = D[1][2, 2] * cosag - D[1][2, 0] * cosamg
    R[1][2, 0] = -D[1][2, 2] * sinag - D[1][2, 0] * sinamg
    R[1][0, 2] = D[1][2, 2] * sinag - D[1][2, 0] * sinamg
    R[1][0, 0] = D[1][2, 2] * cosag + D[1][2, 0] * cosamg
    for l in range(2, ydeg + 1):
        Dl, Dm1, Dm2, Rl = D[l], D[l - 1], D[l - 2], R[l]
        lo, hi = 1 - l, l - 1
        # top row
        Dl[2 * l, 2 * l] = _pmul(Dm1[hi + l - 1, hi + l - 1], [0, 0, 1])
        Dl[2 * l, 0] = _pmul(Dm1[hi + l - 1, -hi + l - 1], [1, 0, 0])
        for m in range(hi, lo - 1, -1):
            x = -np.sqrt((l + m + 1.0) / (l - m)) * Dl[2 * l, m + 1 + l]
            Dl[2 * l, m + l] = np.append(x[1:], [0])
        # upper quarter triangle
        for mp in range(l - 1, -1, -1):
            laux, lbux = l + mp, l - mp
            aux = 1.0 / ((l - 1) * np.sqrt(laux * lbux))
            cux = np.sqrt((laux - 1) * (lbux - 1)) * l
            for m in range(hi, lo - 1, -1):
                lauz, lbuz = l + m, l - m
                fact = aux * (1.0 / np.sqrt(lauz * lbuz))
                a = l * (l - 1)
                b = -(m * mp) / a
                Dl[mp + l, m + l] = _pmul(fact * (2 * l - 1) * a * Dm1[mp + l - 1, m + l - 1],
                                          [b - 1, 0, b + 1])
                if (lbuz != 1) and (lbux != 1):
                    cuz = np.sqrt((lauz - 1) * (lbuz - 1))
                    Dl[mp + l, m + l] -= (fact * cux * cuz) * _pmul(Dm2[mp + l - 2, m + l - 2],
                                                                    [1, 0, 2, 0, 1])
            lo += 1
            hi -= 1
        # reflection / inversion symmetries
        sign, lo, hi = 1, -l, l - 1
        for m in range(l, 0, -1):
            for mp in range(lo, hi + 1):
                Dl[mp + l, m + l] = sign * Dl[m + l, mp + l]
                sign *= -1
            lo += 1
            hi -= 1
        lo = -l
        hi = lo
        for m in range(l - 1, -(l + 1), -1):
            sign = -1
            for mp in range(hi, lo - 1, -1):
                Dl[mp + l, m + l] = sign * Dl[-mp + l, -m + l]
                sign *= -1
            hi += 1
        # complex -> real
        Rl[l, l] = Dl[l, l]
        cosmal, sinmal, sign = c1, s1, -1
        for mp in range(1, l + 1):
            cosmga, sinmga = c3, s3
            aux = rt2 * Dl[l, mp + l]
            Rl[mp + l, l] = aux * cosmal
            Rl[-mp + l, l] = aux * sinmal
            for m in range(1, l + 1):
                aux = rt2 * Dl[m + l, l]
                Rl[l, m + l] = aux * cosmga
                Rl[l, -m + l] = -aux * sinmga
                d1 = Dl[-mp + l, -m + l]
                d2 = sign * Dl[mp + l, -m + l]
                cag = cosmal * cosmga - sinmal * sinmga
                cagm = cosmal * cosmga + sinmal * sinmga
                sag = sinmal * cosmga + cosmal * sinmga
                sagm = sinmal * cosmga - cosmal * sinmga
                Rl[mp + l, m + l] = d1 * cag + d2 * cagm
                Rl[mp + l, -m + l] = -d1 * sag + d2 * sagm
                Rl[-mp + l, m + l] = d1 * sag + d2 * sagm
                Rl[-mp + l, -m + l] = d1 * cag - d2 * cagm
                aux = cosmga * c3 - sinmga * s3
                sinmga = sinmga * c3 + cosmga * s3
                cosmga = aux
            sign *= -1
            aux = cosmal * c1 - sinmal * s1
            sinmal = sinmal * c1 + cosmal * s1
            cosmal = aux
    return R


# ----------------------------------------------------------------------------------------------
# numeric real Wigner x-rotation (host twin of csrc/wigner.cu; ops/include/wigner.h:37-284)
# ----------------------------------------------------------------------------------------------
def rx_numeric(theta, ydeg=YDEG):
    rt2 = math.sqrt(2.0)
    c2, s2 = math.cos(theta), math.sin(theta)
    D = [np.zeros((2 * l + 1, 2 * l + 1)) for l in range(ydeg + 1)]
    R = [np.zeros((2 * l + 1, 2 * l + 1)) for l in range(ydeg + 1)]
    D[0][0, 0] = 1.0
    R[0][0, 0] = 1.0
    d = D[1]
    d[2, 2] = 0.5 * (1.0 + c2)
    d[2, 1] = -s2 / rt2
    d[2, 0] = 0.5 * (1.0 - c2)
    d[1, 2] = -d[2, 1]
    d[1, 1] = d[2, 2] - d[2, 0]
    d[1, 0] = d[2, 1]
    d[0, 2] = d[2, 0]
    d[0, 1] = d[1, 2]
    d[0, 0] = d[2, 2]
    r = R[1]
    r[0, 0] = d[2, 2] - d[2, 0]
    r[0, 1] = -rt2 * d[1, 2]
    r[0, 2] = 0.0
    r[1, 0] = -rt2 * d[2, 1]
    r[1, 1] = d[1, 1]
    r[1, 2] = 0.0
    r[2, 0] = 0.0
    r[2, 1] = 0.0
    r[2, 2] = d[2, 2] + d[2, 0]
    tg = s2 if abs(s2) < 1.0e-14 else (1.0 - c2) / s2
    for l in range(2, ydeg + 1):
        Dl, Dm1, Dm2 = D[l], D[l - 1], D[l - 2]
        lo, hi = 1 - l, l - 1
        Dl[2 * l, 2 * l] = 0.5 * Dm1[hi + l - 1, hi + l - 1] * (1.0 + c2)
        Dl[2 * l, 0] = 0.5 * Dm1[hi + l - 1, -hi + l - 1] * (1.0 - c2)
        for m in range(hi, lo - 1, -1):
            Dl[2 * l, m + l] = -tg * math.sqrt((l + m + 1) / (l - m)) * Dl[2 * l, m + 1 + l]
        al, al1 = l, l - 1
        tal1 = al + al1
        ali = 1.0 / al1
        cosaux = c2 * al * al1
        for mp in range(l - 1, -1, -1):
            laux, lbux = l + mp, l - mp
            aux = ali / math.sqrt(laux * lbux)
            cux = math.sqrt((laux - 1) * (lbux - 1)) * al
            for m in range(hi, lo - 1, -1):
                lauz, lbuz = l + m, l - m
                fact = aux * (1.0 / math.sqrt(lauz * lbuz))
                term = tal1 * (cosaux - float(m * mp)) * Dm1[mp + l - 1, m + l - 1]
                if (lbuz != 1) and (lbux != 1):
                    cuz = math.sqrt((lauz - 1) * (lbuz - 1))
                    term = term - Dm2[mp + l - 2, m + l - 2] * cux * cuz
                Dl[mp + l, m + l] = fact * term
            lo += 1
            hi -= 1
        sign, lo, hi = 1, -l, l - 1
        for m in range(l, 0, -1):
            for mp in range(lo, hi + 1):
                Dl[mp + l, m + l] = sign * Dl[m + l, mp + l]
                sign *= -1
            lo += 1
            hi -= 1
        lo = -l
        hi = lo
        for m in range(l - 1, -(l + 1), -1):
            sign = -1
            for mp in range(hi, lo - 1, -1):
                Dl[mp + l, m + l] = sign * Dl[-mp + l, -m + l]
                sign *= -1
            hi += 1
        Rl = R[l]
        Rl[l, l] = Dl[l, l]
        cosmal, sinmal, sign = 0, -1, -1
        for mp in range(1, l + 1):
            cosmga, sinmga = 0, 1
            Rl[mp + l, l] = rt2 * Dl[l, mp + l] * cosmal
            Rl[-mp + l, l] = rt2 * Dl[l, mp + l] * sinmal
            for m in range(1, l + 1):
                d1 = Dl[-mp + l, -m + l]
                d2 = sign * Dl[mp + l, -m + l]
                cag = cosmal * cosmga - sinmal * sinmga
                cagm = cosmal * cosmga + sinmal * sinmga
                sag = sinmal * cosmga + cosmal * sinmga
                sagm = sinmal * cosmga - cosmal * sinmga
                Rl[l, m + l] = rt2 * Dl[m + l, l] * cosmga
                Rl[l, -m + l] = -rt2 * Dl[m + l, l] * sinmga
                Rl[mp + l, m + l] = d1 * cag + d2 * cagm
                Rl[mp + l, -m + l] = -d1 * sag + d2 * sagm
                Rl[-mp + l, m + l] = d1 * sag + d2 * sagm
                Rl[-mp + l, -m + l] = d1 * cag - d2 * cagm
                cosmga, sinmga = -sinmga, cosmga
            sign *= -1
            cosmal, sinmal = sinmal, -cosmal
    return np.concatenate([x.reshape(-1) for x in R])


def rx90_nonzeros(rx90):
    """Structural non-zeros of the packed Rx(pi/2): a rotation by pi/2 about x commutes with the
    reflections x -> -x and (y, z) -> (z, -y) ..., so each (2l+1)^2 block is a 4-way permuted block
    diagonal and only 1372 of the 5456 entries are non-zero (the rest are <= 6e-16, against
    >= 6e-5 for the true entries).  Returned in the order the unrolled design-matrix kernel
    consumes them (csrc/gen_design.py): degree l, output column m, then source row m'.
    Returns (values, [(l, mp_index, m_index), ...])."""
    vals, idx = [], []
    for l in range(YDEG + 1):
        w = 2 * l + 1
        blk = rx90[nwig(l - 1):nwig(l)].reshape(w, w)
        for j in range(w):
            for mp in range(w):
                if abs(blk[mp, j]) > 1e-12:
                    vals.append(blk[mp, j])
                    idx.append((l, mp, j))
    return np.array(vals), idx


def _blockdiag(packed):
    M = np.zeros((N, N))
    for l in range(YDEG + 1):
        M[l * l:(l + 1) ** 2, l * l:(l + 1) ** 2] = packed[nwig(l - 1):nwig(l)].reshape(
            2 * l + 1, 2 * l + 1)
    return M


def _matrix_sqrt(Q, neig):
    """math.py:121-139 with the numpy driver (ops/eigh/eigh.py:11-20)."""
    w, U = np.linalg.eigh(Q)
    w = w[-neig:]
    U = U[:, -neig:]
    with np.errstate(invalid="ignore"):
        sw = np.where(w > 1e-15, np.sqrt(w), 0.0)
    return U @ np.diag(sw)


def spot_operator(ydeg):
    """size.py:10-43: ``Bp = S A`` (ydeg + 1, 1000), the smoothed least-squares Legendre fit of a spot
    profile sampled on ``theta = linspace(0, pi, 1000)``.  The fit degree enters ``A`` (the Legendre
    polynomials are not orthogonal on the grid), so a lower ``ydeg`` has its own operator."""
    theta = np.linspace(0, np.pi, SPTS)
    cost = np.cos(theta)
    B = np.hstack([np.sqrt(2 * l + 1) * _legendre(l)(cost).reshape(-1, 1) for l in range(ydeg + 1)])
    A = np.linalg.solve(B.T @ B + 1e-9 * np.eye(ydeg + 1), B.T)
    ll = np.arange(ydeg + 1)
    S = np.exp(-0.5 * ll * (ll + 1) * 0.075 ** 2)
    return S[:, None] * A


_CACHE = {}
LONGITUDE_BASES = ("pinned", "host")


def build_tables(longitude_basis="pinned", use_pinned_longitude=None):
    """Returns (blob float64[_TOTAL], offsets dict).

    ``longitude_basis`` selects the hyperparameter-independent longitude eigenvector table
    ``U_lon = matrix_sqrt(Q_lon)`` (longitude.py:9-49 -> integrals.py:116-124 -> math.py:121-139):

    * ``"pinned"`` (default): the table shipped in ``data/longitude_U_ydeg15.npy`` (regenerate it
      with ``scripts/gen_longitude_U.py``; provenance in the ``.json`` next to it).  ``Q_lon`` has
      eigenvalues at the 1e-15 clip whose eigenvectors depend on the LAPACK build / thread count, and
      the reference's lnlike moves by up to 3e-6 with them; the pinned table makes every host
      reproduce the values of the build container that generated the golden fixtures.
    * ``"host"``: ``numpy.linalg.eigh`` on THIS host, exactly what the reference (and the CPU
      oracle) computes in this process -- the basis to use when comparing against a reference run on
      the same machine.
    """
    if use_pinned_longitude is not None:   # round-1 spelling
        longitude_basis = "pinned" if use_pinned_longitude else "host"
    if longitude_basis not in LONGITUDE_BASES:
        raise ValueError("longitude_basis must be one of %r" % (LONGITUDE_BASES,))
    if ("blob", longitude_basis) in _CACHE:
        return _CACHE[("blob", longitude_basis)], offsets()
    off = offsets()
    if "common" in _CACHE:
        blob = _CACHE["common"].copy()
        _put_longitude(blob, off, longitude_basis)
        _CACHE[("blob", longitude_basis)] = blob
        return blob, off
    blob = np.zeros(off["_TOTAL"])

    def put(name, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        cnt = dict(LAYOUT)[name]
        assert arr.size == cnt, (name, arr.size, cnt)
        blob[off[name]:off[name] + cnt] = arr

    l_of, m_of = _lm()

    # ---- spot profile operator (size.py:10-43)
    theta = np.linspace(0, np.pi, SPTS)
    cost = np.cos(theta)
    B = np.hstack([np.sqrt(2 * l + 1) * _legendre(l)(cost).reshape(-1, 1) for l in range(YDEG + 1)])
    A = np.linalg.solve(B.T @ B + 1e-9 * np.eye(YDEG + 1), B.T)
    ll = np.arange(YDEG + 1)
    S = np.exp(-0.5 * ll * (ll + 1) * 0.075 ** 2)
    put("THETA", theta)
    put("BP", S[:, None] * A)

    # ---- latitude (Euler angles of latitude.py:203-205)
    R_lat = wigner_poly(YDEG, 0, 1, 0, -1)
    # range of the rank-31 second-moment matrix: monomials c^j s^(2l-j) promoted to degree 30
    V = np.zeros((N, NEIG))
    for l in range(YDEG + 1):
        for k in range(2 * l + 1):
            for t in range(YDEG - l + 1):
                V[l * l + k, k + 2 * t] = math.comb(YDEG - l, t)
    Z = np.linalg.svd(V, full_matrices=False)[0]
    Zp = np.zeros((N, 32))
    Zp[:, :NEIG] = Z
    H = np.zeros((N, 32))
    R0 = np.zeros(NWIG)
    for l in range(YDEG + 1):
        r0 = R_lat[l][:, l, :]                      # [m', k]
        R0[nwig(l - 1):nwig(l)] = r0.reshape(-1)
        H[l * l:(l + 1) ** 2, :NEIG] = (r0.astype(np.longdouble)
                                        @ Z[l * l:(l + 1) ** 2].astype(np.longdouble)).astype(
            np.float64)
    # binomial factors of the double sum term(2a, 2b) (latitude.h:117-143), generated with the
    # reference's own ratio recurrences (same IEEE operations, same order) so that the kernel can
    # replace its FP64 divisions by table look-ups without changing a single bit
    fac_tab = []
    fac_off = np.zeros(31 * 31 + 1)
    for a2 in range(31):
        for b2 in range(31):
            fac_off[a2 * 31 + b2] = len(fac_tab)
            if a2 + b2 > 30:
                continue
            fac1 = 1.0
            for k1 in range(a2 + 1):
                fac2 = fac1
                for k2 in range(b2 + 1):
                    fac_tab.append(fac2)
                    fac2 *= (k2 - b2) / (k2 + 1.0)
                fac1 *= (a2 - k1) / (k1 + 1.0)
    fac_off[31 * 31] = len(fac_tab)
    put("LAT_FAC", np.array(fac_tab))
    put("LAT_FACOFF", fac_off)
    put("LAT_Z", Zp)
    put("LAT_H", H)
    put("LAT_R0", R0)

    # ---- longitude (longitude.py:9-49, integrals.py:116-124): filled per basis by _put_longitude
    n4 = 4 * YDEG + 1

    # ---- flux integrals (flux.py:107-179)
    def _G(j, i):
        return 2 * _gamma(1 + 0.5 * i) * _gamma(1 + 0.5 * j) / _gamma(0.5 * (4 + i + j)) - (
            2 ** (1 - 0.5 * i) / (2 + i)) * _hyp2f1(1 + 0.5 * i, -0.5 * j, 2 + 0.5 * i, 0.5)

    G = np.array([[_G(i, j) for i in range(n4)] for j in range(n4)])
    wnp = np.zeros(NWIG)
    for l in range(YDEG + 1):
        m = np.arange(-l, l + 1)
        wnp[nwig(l - 1):nwig(l)] = (R_lat[l] @ G[l - m, l + m]).reshape(-1)
    Qt = np.empty((NEIG, NEIG, NEIG, N))
    for l1 in range(YDEG + 1):
        k = np.arange(l1 ** 2, (l1 + 1) ** 2)
        k0 = np.arange(2 * l1 + 1).reshape(-1, 1)
        for p in range(N):
            l2 = int(np.floor(np.sqrt(p)))
            j = np.arange(l2 ** 2, (l2 + 1) ** 2)
            j0 = np.arange(2 * l2 + 1).reshape(1, -1)
            Lm = R_lat[l1][l1, k - l1 ** 2] @ G[k0 + j0, 2 * l1 - k0 + 2 * l2 - j0]
            Rm = R_lat[l2][j - l2 ** 2, p - l2 ** 2].T
            Qt[l1, : 2 * l1 + 1, : 2 * l2 + 1, p] = Lm @ Rm
    Wnp = np.empty((N, N))
    for l1 in range(YDEG + 1):
        i = np.arange(l1 ** 2, (l1 + 1) ** 2)
        for l2 in range(YDEG + 1):
            j = np.arange(l2 ** 2, (l2 + 1) ** 2)
            Wnp[i.reshape(-1, 1), j.reshape(1, -1)] = Qt[l1, : 2 * l1 + 1, l2, j].T
    put("FLUX_WNP", wnp)
    put("FLUX_W", Wnp)

    # ---- Rx(pi/2) and the folded quadratic forms
    #   a_m = sum_{i in group m} sum_j W_ij Ez_ij,  Ez = Rx^T (Sigma + mu mu^T) Rx
    #       = <Omega_m, Sigma> + (mu-part),  Omega_m[p, q] = sum_{i in group m} Rx[p, i] (Wnp Rx^T)[i, q]
    rx90 = rx_numeric(0.5 * np.pi)
    put("RX90", rx90)
    put("RX90_NZ", rx90_nonzeros(rx90)[0])
    Rx = _blockdiag(rx90).astype(np.longdouble)
    WR = Wnp.astype(np.longdouble) @ Rx.T
    #   b_m = sum_{i in group m} sgn(m_i) sum_j W_ij Ez_{i jbar}   (the sine lane of wigner.h:440-458;
    #       zero for an exactly symmetric Ez, but the reference carries its rounding-level value)
    bar = l_of * l_of + l_of - m_of
    WRb = Wnp[:, bar].astype(np.longdouble) @ Rx.T
    Om = np.zeros((31, N, N))
    for mm in range(16):
        sel = np.abs(m_of) == mm
        Om[mm] = (Rx[:, sel] @ WR[sel, :]).astype(np.float64)
        if mm > 0:
            sg = np.sign(m_of[sel]).astype(np.longdouble)
            Om[15 + mm] = ((Rx[:, sel] * sg[None, :]) @ WRb[sel, :]).astype(np.float64)
    put("FLUX_OMEGA", Om)

    # ---- Gauss-Legendre rule on [0, 1] (32 nodes: exact to degree 63)
    xg, wg = np.polynomial.legendre.leggauss(32)
    put("GL_X", 0.5 * (xg + 1.0))
    put("GL_W", 0.5 * wg)

    # ---- jitter (contrast.py:26-32)
    lam = np.ones(N) * 1e-12
    lam[15 ** 2:] = 1e-9
    put("LAMBDA", lam)

    _CACHE["common"] = blob.copy()
    _put_longitude(blob, off, longitude_basis)
    _CACHE[("blob", longitude_basis)] = blob
    return blob, off


def longitude_qQ():
    """longitude.py:22-49: first / second moment integrals of the (uniform) longitude prior."""
    l_of, m_of = _lm()
    n4 = 4 * YDEG + 1
    term = np.zeros((n4, n4))
    for i in range(n4):
        for j in range(0, n4, 2):
            term[i, j] = _gamma(0.5 * (i + 1)) * _gamma(0.5 * (j + 1)) / _gamma(0.5 * (2 + i + j))
    term /= np.pi
    jj = m_of + l_of
    ii = l_of - m_of
    q_lon = term[jj, ii]
    Q_lon = term[jj[:, None] + jj[None, :], ii[:, None] + ii[None, :]]
    return q_lon, Q_lon


PINNED_LONGITUDE = os.path.join(HERE, "data", "longitude_U_ydeg15.npy")


def longitude_U(basis):
    """``U_lon (256, 31)`` for the requested basis (see build_tables)."""
    if basis == "pinned":
        if not os.path.exists(PINNED_LONGITUDE):
            raise RuntimeError("%s is missing: regenerate it with scripts/gen_longitude_U.py"
                               % PINNED_LONGITUDE)
        U = np.load(PINNED_LONGITUDE)
        if U.shape != (N, NEIG):
            raise RuntimeError("pinned longitude table has the wrong shape %r" % (U.shape,))
        return U
    return _matrix_sqrt(longitude_qQ()[1], NEIG)


def _put_longitude(blob, off, basis):
    if "R_lon" not in _CACHE:
        _CACHE["R_lon"] = wigner_poly(YDEG, 1, 0, 1, 0)
    R_lon = _CACHE["R_lon"]
    q_lon, _ = longitude_qQ()
    U_lon = longitude_U(basis)
    T1 = np.zeros(NWIG)
    TL = np.zeros(NEIG * NWIG)
    pos = 0
    for l in range(YDEG + 1):
        w = 2 * l + 1
        T1[nwig(l - 1):nwig(l)] = np.dot(R_lon[l], q_lon[l * l:(l + 1) ** 2]).reshape(-1)
        # T[l][m', e2, m] = sum_k R[l][m', m, k] U[l^2 + k, e2]
        Tl = np.swapaxes(np.dot(R_lon[l], U_lon[l * l:(l + 1) ** 2]), 1, 2)
        TL[pos:pos + w * NEIG * w] = Tl.reshape(-1)
        pos += w * NEIG * w
    blob[off["LON_T1"]:off["LON_T1"] + NWIG] = T1
    blob[off["LON_T"]:off["LON_T"] + NEIG * NWIG] = TL
    _CACHE[("U_lon", basis)] = U_lon
