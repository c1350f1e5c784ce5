"""
TEST INFRASTRUCTURE ONLY -- the CPU oracle for the batched log-likelihood hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` leg may import this module; the product package ``starry_process_b200`` never does.

This is a plain NumPy/SciPy restatement of the reference's *Python glue* (Theano graph replaced
by eager arrays), following the reference file by file; the native numerics come from either

* ``oracle/liboracle_native.so``  -- the plain-C restatement in ``oracle_native.c`` (default), or
* ``oracle/_ref/libspref_y*_u*.so`` -- the reference's own C++ headers compiled in place
  (``native="ref"``), used to validate the restatement and as the "reference" CPU baseline.

PARITY PIN.  The reference's own tests hold no absolute lnlike values.
The oracle is therefore pinned against outputs of the reference ITSELF run in the build
container: ``oracle/theano_stub.py`` lets the unmodified reference package execute eagerly, and
``oracle/gen_golden.py`` stores its outputs under ``tests/golden/``.  ``tests/test_oracle_*.py``
check this module against those fixtures (everywhere) and against the live reference (when
``/root/reference`` exists), and against the reference's one golden artefact,
``app/data/A_F15-300.npz`` (sub-sampled into ``tests/golden/design_matrix_AF15.npz``).

Citations are ``file:line`` under ``/root/reference/starry_process/``.
"""
import ctypes
import os
import threading

import numpy as np
import scipy.linalg
from scipy.special import gamma as _gamma
from scipy.special import hyp2f1 as _sp_hyp2f1
from scipy.special import legendre as _legendre

HERE = os.path.dirname(os.path.abspath(__file__))

# defaults.py:4-35
DEFAULTS = dict(
    ydeg=15, udeg=2, r=20.0, dr=None, a=0.40, b=0.27, c=0.1, n=10.0, p=1.0, i=60.0,
    normalized=True, normalization_order=20, normalization_zmax=0.023,
    marginalize_over_inclination=True, baseline_mean=0.0, baseline_var=0.0,
    eps=1e-8, epsy=1e-12, epsy15=1e-9, covpts=300, log_alpha_max=10, log_beta_max=10,
    abmin=1e-12, sigma_max=45.0,
)


# ----------------------------------------------------------------------------------------------
# native back ends
# ----------------------------------------------------------------------------------------------
def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class _NativeC(object):
    """oracle_native.c (plain-C restatement)."""

    kind = "port"

    def __init__(self):
        path = os.path.join(HERE, "liboracle_native.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle_native.so missing: run `make -C oracle native`")
        lib = ctypes.CDLL(path)
        D, P, I = ctypes.c_double, ctypes.c_void_p, ctypes.c_int
        lib.orc_hyp2f1.argtypes = [D, D, D, D]
        lib.orc_hyp2f1.restype = D
        lib.orc_latitude_integrals.argtypes = [I, D, D, P, P]
        lib.orc_Rx.argtypes = [I, D, P]
        lib.orc_tensordotRz.argtypes = [I, P, P, I, P]
        lib.orc_special_tensordotRz.argtypes = [I, P, P, P, I, P]
        lib.orc_rTA1.argtypes = [I, P]
        lib.orc_rTA1L.argtypes = [I, I, P, P]
        for nm in ("orc_latitude_integrals", "orc_Rx", "orc_tensordotRz",
                   "orc_special_tensordotRz", "orc_rTA1", "orc_rTA1L"):
            getattr(lib, nm).restype = None
        self.lib = lib

    def hyp2f1(self, a, b, c, z):
        return self.lib.orc_hyp2f1(a, b, c, z)

    def latitude(self, ydeg, udeg, alpha, beta):
        N = (ydeg + 1) ** 2
        q = np.empty(N)
        Q = np.empty((N, N))
        self.lib.orc_latitude_integrals(ydeg, float(alpha), float(beta), _ptr(q), _ptr(Q))
        return q, Q

    def Rx(self, ydeg, udeg, theta):
        R = np.empty(nwig(ydeg))
        self.lib.orc_Rx(ydeg, float(theta), _ptr(R))
        return R

    def tensordotRz(self, ydeg, udeg, M, theta):
        M = np.ascontiguousarray(M, dtype=np.float64)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        f = np.empty_like(M)
        self.lib.orc_tensordotRz(ydeg, _ptr(M), _ptr(theta), theta.shape[0], _ptr(f))
        return f

    def special_tensordotRz(self, ydeg, udeg, T, M, theta):
        T = np.ascontiguousarray(T, dtype=np.float64)
        M = np.ascontiguousarray(M, dtype=np.float64)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        f = np.empty(theta.shape[0])
        self.lib.orc_special_tensordotRz(ydeg, _ptr(T), _ptr(M), _ptr(theta), theta.shape[0],
                                         _ptr(f))
        return f

    def rTA1(self, ydeg, udeg):
        f = np.empty((ydeg + 1) ** 2)
        self.lib.orc_rTA1(ydeg, _ptr(f))
        return f

    def rTA1L(self, ydeg, udeg, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        f = np.empty((ydeg + 1) ** 2)
        self.lib.orc_rTA1L(ydeg, udeg, _ptr(u), _ptr(f))
        return f


class _NativeRef(object):
    """oracle/_ref: the reference's own headers behind ref_shim.cc (compile-time ydeg/udeg)."""

    kind = "reference"

    def __init__(self):
        self._libs = {}

    @staticmethod
    def available(ydeg=15, udeg=2):
        return os.path.exists(os.path.join(HERE, "_ref", "libspref_y%d_u%d.so" % (ydeg, udeg)))

    def _lib(self, ydeg, udeg):
        key = (ydeg, udeg)
        if key not in self._libs:
            path = os.path.join(HERE, "_ref", "libspref_y%d_u%d.so" % key)
            if not os.path.exists(path):
                raise RuntimeError("%s missing: run `make -C oracle ref` where /root/reference "
                                   "exists" % path)
            lib = ctypes.CDLL(path)
            D, P, I = ctypes.c_double, ctypes.c_void_p, ctypes.c_int
            lib.ref_Rx.argtypes = [D, P, P]
            lib.ref_tensordotRz.argtypes = [P, P, I, P]
            lib.ref_special_tensordotRz.argtypes = [P, P, P, I, P]
            lib.ref_rTA1.argtypes = [P]
            lib.ref_rTA1L.argtypes = [P, P]
            lib.ref_latitude.argtypes = [D, D, P, P]
            lib.ref_hyp2f1.argtypes = [D, D, D, D]
            lib.ref_hyp2f1.restype = D
            for nm in ("ref_Rx", "ref_tensordotRz", "ref_special_tensordotRz", "ref_rTA1",
                       "ref_rTA1L", "ref_latitude"):
                getattr(lib, nm).restype = None
            self._libs[key] = lib
        return self._libs[key]

    def hyp2f1(self, a, b, c, z):
        return self._lib(15, 2).ref_hyp2f1(a, b, c, z)

    def latitude(self, ydeg, udeg, alpha, beta):
        N = (ydeg + 1) ** 2
        q = np.empty(N)
        Q = np.empty((N, N))
        self._lib(ydeg, udeg).ref_latitude(float(alpha), float(beta), _ptr(q), _ptr(Q))
        return q, Q

    def Rx(self, ydeg, udeg, theta):
        R = np.empty(nwig(ydeg))
        dR = np.empty(nwig(ydeg))
        self._lib(ydeg, udeg).ref_Rx(float(theta), _ptr(R), _ptr(dR))
        return R

    def tensordotRz(self, ydeg, udeg, M, theta):
        M = np.ascontiguousarray(M, dtype=np.float64)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        f = np.empty_like(M)
        self._lib(ydeg, udeg).ref_tensordotRz(_ptr(M), _ptr(theta), theta.shape[0], _ptr(f))
        return f

    def special_tensordotRz(self, ydeg, udeg, T, M, theta):
        T = np.ascontiguousarray(T, dtype=np.float64)
        M = np.ascontiguousarray(M, dtype=np.float64)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        f = np.empty(theta.shape[0])
        self._lib(ydeg, udeg).ref_special_tensordotRz(_ptr(T), _ptr(M), _ptr(theta),
                                                      theta.shape[0], _ptr(f))
        return f

    def rTA1(self, ydeg, udeg):
        f = np.empty((ydeg + 1) ** 2)
        self._lib(ydeg, udeg).ref_rTA1(_ptr(f))
        return f

    def rTA1L(self, ydeg, udeg, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        f = np.empty((ydeg + 1) ** 2)
        self._lib(ydeg, udeg).ref_rTA1L(_ptr(u), _ptr(f))
        return f


_NATIVE = {}
_LOCK = threading.Lock()


def get_native(kind="port"):
    with _LOCK:
        if kind not in _NATIVE:
            _NATIVE[kind] = _NativeC() if kind == "port" else _NativeRef()
        return _NATIVE[kind]


def ref_available(ydeg=15, udeg=2):
    return _NativeRef.available(ydeg, udeg)


def nwig(l):
    """wigner.h:22-24 / flux.py:77."""
    return ((l + 1) * (2 * l + 1) * (2 * l + 3)) // 3


# ----------------------------------------------------------------------------------------------
# wigner.py:155-372  polynomial-coefficient Wigner matrices
# ----------------------------------------------------------------------------------------------
def _prod(x1, x2):
    """wigner.py:155-164 -- polynomial product.  result[m + n] += x1[m] * x2[n] with the
    reference's accumulation order (m outer, ascending); separate multiply and add ufuncs, so no
    fused multiply-add on any host."""
    x1 = np.asarray(x1, dtype=float)
    x2 = np.asarray(x2, dtype=float)
    result = np.zeros(len(x1) + len(x2) - 1)
    for m in range(len(x1)):
        result[m:m + len(x2)] += x1[m] * x2
    return result


def _poly_dlmn(l, s1, c1, s3, c3, D, R):
    """wigner.py:192-292."""
    iinf = 1 - l
    isup = -iinf
    # first row by recurrence, wigner.py:199-206
    D[l][2 * l, 2 * l] = _prod(D[l - 1][isup + l - 1, isup + l - 1], [0, 0, 1])
    D[l][2 * l, 0] = _prod(D[l - 1][isup + l - 1, -isup + l - 1], [1, 0, 0])
    for m in range(isup, iinf - 1, -1):
        x = -np.sqrt((l + m + 1.0) / (l - m)) * D[l][2 * l, m + 1 + l]
        D[l][2 * l, m + l] = np.append(x[1:], [0])
    # upper quarter triangle, wigner.py:210-235
    for mp in range(l - 1, -1, -1):
        laux = l + mp
        lbux = l - mp
        aux = 1.0 / ((l - 1) * np.sqrt(laux * lbux))
        cux = np.sqrt((laux - 1) * (lbux - 1)) * l
        for m in range(isup, iinf - 1, -1):
            lauz = l + m
            lbuz = l - m
            auz = 1.0 / np.sqrt(lauz * lbuz)
            fact = aux * auz
            a = l * (l - 1)
            b = -(m * mp) / a
            D[l][mp + l, m + l] = _prod(
                fact * (2 * l - 1) * a * D[l - 1][mp + l - 1, m + l - 1], [b - 1, 0, b + 1]
            )
            if (lbuz != 1) and (lbux != 1):
                cuz = np.sqrt(((lauz - 1) * (lbuz - 1)))
                D[l][mp + l, m + l] -= (fact * cux * cuz) * _prod(
                    D[l - 2][mp + l - 2, m + l - 2], [1, 0, 2, 0, 1]
                )
        iinf += 1
        isup -= 1
    # reflection, wigner.py:243-252
    sign = 1
    iinf = -l
    isup = l - 1
    for m in range(l, 0, -1):
        for mp in range(iinf, isup + 1):
            D[l][mp + l, m + l] = sign * D[l][m + l, mp + l]
            sign *= -1
        iinf += 1
        isup -= 1
    # inversion, wigner.py:254-262
    iinf = -l
    isup = iinf
    for m in range(l - 1, -(l + 1), -1):
        sign = -1
        for mp in range(isup, iinf - 1, -1):
            D[l][mp + l, m + l] = sign * D[l][-mp + l, -m + l]
            sign *= -1
        isup += 1
    # real from complex, wigner.py:264-292
    R[l][l, l] = D[l][l, l]
    cosmal, sinmal = c1, s1
    sign = -1
    root_two = np.sqrt(2.0)
    for mp in range(1, l + 1):
        cosmga, sinmga = c3, s3
        aux = root_two * D[l][0 + l, mp + l]
        R[l][mp + l, 0 + l] = aux * cosmal
        R[l][-mp + l, 0 + l] = aux * sinmal
        for m in range(1, l + 1):
            aux = root_two * D[l][m + l, 0 + l]
            R[l][l, m + l] = aux * cosmga
            R[l][l, -m + l] = -aux * sinmga
            d1 = D[l][-mp + l, -m + l]
            d2 = sign * D[l][mp + l, -m + l]
            cosag = cosmal * cosmga - sinmal * sinmga
            cosagm = cosmal * cosmga + sinmal * sinmga
            sinag = sinmal * cosmga + cosmal * sinmga
            sinagm = sinmal * cosmga - cosmal * sinmga
            R[l][mp + l, m + l] = d1 * cosag + d2 * cosagm
            R[l][mp + l, -m + l] = -d1 * sinag + d2 * sinagm
            R[l][-mp + l, m + l] = d1 * sinag + d2 * sinagm
            R[l][-mp + l, -m + l] = d1 * cosag - d2 * cosagm
            aux = cosmga * c3 - sinmga * s3
            sinmga = sinmga * c3 + cosmga * s3
            cosmga = aux
        sign *= -1
        aux = cosmal * c1 - sinmal * s1
        sinmal = sinmal * c1 + cosmal * s1
        cosmal = aux


_POLY_R_CACHE = {}


def wigner_poly_R(ydeg, cos_alpha=0, sin_alpha=1, cos_gamma=0, sin_gamma=-1):
    """wigner.py:295-372 with ``phi=None``: R[l][m', m, k], k = power of cos(phi/2)."""
    key = (ydeg, cos_alpha, sin_alpha, cos_gamma, sin_gamma)
    if key in _POLY_R_CACHE:
        return _POLY_R_CACHE[key]
    c1, s1, c3, s3 = cos_alpha, sin_alpha, cos_gamma, sin_gamma
    root_two = np.sqrt(2.0)
    D = [np.nan * np.ones((2 * l + 1,) * 3) for l in range(ydeg + 1)]
    R = [np.nan * np.ones((2 * l + 1,) * 3) for l in range(ydeg + 1)]
    D[0][0, 0] = [1]
    R[0][0, 0] = [1]
    D[1][2, 2] = [0, 0, 1]
    D[1][2, 1] = [0, -root_two, 0]
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
    R[1][2, 1] = root_two * D[1][1, 2] * c1
    R[1][0, 1] = root_two * D[1][1, 2] * s1
    R[1][1, 2] = root_two * D[1][2, 1] * c3
    R[1][1, 0] = -root_two * D[1][2, 1] * s3
    R[1][2, 2] = D[1][2, 2] * cosag - D[1][2, 0] * cosamg
    R[1][2, 0] = -D[1][2, 2] * sinag - D[1][2, 0] * sinamg
    R[1][0, 2] = D[1][2, 2] * sinag - D[1][2, 0] * sinamg
    R[1][0, 0] = D[1][2, 2] * cosag + D[1][2, 0] * cosamg
    for l in range(2, ydeg + 1):
        _poly_dlmn(l, s1, c1, s3, c3, D, R)
    _POLY_R_CACHE[key] = R
    return R


# ----------------------------------------------------------------------------------------------
# math.py:121-139 + ops/eigh/eigh.py:11-20
# ----------------------------------------------------------------------------------------------
# The reference supports two eigensolver drivers (ops/eigh/eigh.py:11-26, defaults.py:24
# `driver="numpy"`): numpy.linalg.eigh (LAPACK dsyevd) and scipy.linalg.eigh with an index subset
# (dsyevr).  Tests flip this switch to measure how far the REFERENCE's own lnlike moves between its
# two drivers on the same inputs -- its reproducibility floor, which bounds any third-party parity.
EIGH_DRIVER = "numpy"
# Probe only (reference value: 1.0): scales the `w > 1e-15` clip of math.py:134-136.  Noise-level
# eigenvalues of the rank-deficient moment matrices sit right at that clip; whether one survives
# depends on the LAPACK build / CPU, and each survivor moves cov_ylm by ~1e-9.
CLIP_SCALE = 1.0


def matrix_sqrt(Q, neig=None, mindiff=1e-15):
    N = Q.shape[0]
    neig = N if neig is None else neig
    try:
        if EIGH_DRIVER == "scipy":
            w, U = scipy.linalg.eigh(Q, subset_by_index=(N - neig, N - 1))
        else:
            w, U = np.linalg.eigh(Q)
    except np.linalg.LinAlgError:
        return np.full((N, neig), np.nan)
    w = np.ascontiguousarray(w[-neig:])
    U = np.ascontiguousarray(U[:, -neig:])
    with np.errstate(invalid="ignore"):
        sqrtw = np.where(w > mindiff * CLIP_SCALE, np.sqrt(w), 0.0)
    return U @ np.diag(sqrtw)


def reference_noise_floor(fn, ntrials=3):
    """Relative spread of ``fn(**extra_kwargs_for_OracleProcess)`` (a scalar or array of lnlike
    values) between the reference's two eigensolver drivers and under ``ntrials`` one-ulp
    perturbations of the latitude moment matrix: how reproducible the REFERENCE's own number is."""
    global EIGH_DRIVER, CLIP_SCALE
    base = np.asarray(fn(), dtype=float)
    dev = np.zeros_like(base)
    old = EIGH_DRIVER
    try:
        EIGH_DRIVER = "scipy" if old == "numpy" else "numpy"
        dev = np.maximum(dev, np.abs(np.asarray(fn(), dtype=float) - base))
    finally:
        EIGH_DRIVER = old
    try:
        for CLIP_SCALE in (0.25, 4.0):   # toggles modes within a factor 4 of the clip
            dev = np.maximum(dev, np.abs(np.asarray(fn(), dtype=float) - base))
    finally:
        CLIP_SCALE = 1.0
    for k in range(ntrials):
        dev = np.maximum(dev, np.abs(np.asarray(fn(q_ulp_noise_seed=k), dtype=float) - base))
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(np.isfinite(base), dev / np.abs(base), 0.0)


def cho_factor(A):
    """math.py:75-94: lower Cholesky, NaN on failure."""
    try:
        return scipy.linalg.cholesky(A, lower=True)
    except (scipy.linalg.LinAlgError, ValueError):
        return np.full(A.shape, np.nan)


def cho_solve(L, b):
    """math.py:20-38, 97-100."""
    if np.any(np.isnan(L)) or np.any(np.isnan(b)):
        return np.full(b.shape, np.nan)
    y = scipy.linalg.solve_triangular(L, b, lower=True)
    return scipy.linalg.solve_triangular(L.T, y, lower=False)


def check_bounds(name, x, lower=-np.inf, upper=np.inf, tol=1e-6):
    """ops/exceptions.py:30-48."""
    x = np.asarray(x, dtype=float)
    if np.any((x < lower - tol) | (x > upper + tol)):
        if np.any(x < lower - tol):
            value, sign, bound = np.atleast_1d(x)[np.atleast_1d(x < lower - tol)][0], "<=", lower
        else:
            value, sign, bound = np.atleast_1d(x)[np.atleast_1d(x > upper + tol)][0], ">=", upper
        raise ValueError("%s out of bounds: %f %s %f" % (name, value, sign, bound))
    return x


# ----------------------------------------------------------------------------------------------
# latitude.py:14-77, 80-168
# ----------------------------------------------------------------------------------------------
def gauss2beta(mu, sigma, log_alpha_max=10, log_beta_max=10):
    is_vector = hasattr(mu, "__len__")
    m = np.atleast_1d(np.asarray(mu, dtype=float)) * np.pi / 180
    v = (np.atleast_1d(np.asarray(sigma, dtype=float)) * np.pi / 180) ** 2
    c1 = np.cos(m)
    c2 = np.cos(2 * m)
    c3 = np.cos(3 * m)
    term = 1.0 / (16 * v * np.cos(0.5 * m) ** 4)
    alpha = (2 + 4 * v + (3 + 8 * v) * c1 + 2 * c2 + c3) * term
    beta = (c1 + 2 * v * (3 + c2) - c3) * term
    a = np.log(alpha) / log_alpha_max
    b = np.maximum(0.0, (np.log(beta) - np.log(0.5)) / (log_beta_max - np.log(0.5)))
    if is_vector:
        return a, b
    return a[0], b[0]


def beta2gauss(a, b, log_alpha_max=10, log_beta_max=10):
    is_vector = hasattr(a, "__len__")
    alpha = np.atleast_1d(np.exp(np.asarray(a, dtype=float) * log_alpha_max))
    beta = np.atleast_1d(np.exp(np.log(0.5) + np.asarray(b, dtype=float)
                                * (log_beta_max - np.log(0.5))))
    with np.errstate(invalid="ignore", divide="ignore"):
        term = 4 * alpha ** 2 - 8 * alpha - 6 * beta + 4 * alpha * beta + beta ** 2 + 5
        mu = 2 * np.arctan(np.sqrt(2 * alpha + beta - 2 - np.sqrt(term)))
        term = (1 - alpha + beta + (beta - 1) * np.cos(mu) + (alpha - 1) / np.cos(mu) ** 2)
        sigma = np.sin(mu) / np.sqrt(term)
    mu[(alpha <= 1) | (beta <= 0.5)] = np.nan
    sigma[(alpha <= 1) | (beta <= 0.5)] = np.nan
    if is_vector:
        return mu / (np.pi / 180), sigma / (np.pi / 180)
    return mu[0] / (np.pi / 180), sigma[0] / (np.pi / 180)


# ----------------------------------------------------------------------------------------------
# Hyperparameter-independent constants
# ----------------------------------------------------------------------------------------------
_CONST_CACHE = {}


def spot_Bp(ydeg, spts=1000, eps4=1e-9, smoothing=0.075):
    """size.py:10-43."""
    key = ("Bp", ydeg, spts, eps4, smoothing)
    if key not in _CONST_CACHE:
        theta = np.linspace(0, np.pi, spts)
        cost = np.cos(theta)
        B = np.hstack([np.sqrt(2 * l + 1) * _legendre(l)(cost).reshape(-1, 1)
                       for l in range(ydeg + 1)])
        A = np.linalg.solve(B.T @ B + eps4 * np.eye(ydeg + 1), B.T)
        l = np.arange(ydeg + 1)
        i = l * (l + 1)
        S = np.exp(-0.5 * i * smoothing ** 2)
        _CONST_CACHE[key] = (theta, S[:, None] * A, i)
    return _CONST_CACHE[key]


def longitude_qQ(ydeg):
    """longitude.py:22-49."""
    key = ("lonqQ", ydeg)
    if key not in _CONST_CACHE:
        n = 4 * ydeg + 1
        N = (ydeg + 1) ** 2
        term = np.zeros((n, n))
        for i in range(n):
            for j in range(0, n, 2):
                term[i, j] = (_gamma(0.5 * (i + 1)) * _gamma(0.5 * (j + 1))
                              / _gamma(0.5 * (2 + i + j)))
        term /= np.pi
        l = np.concatenate([np.full(2 * ll + 1, ll) for ll in range(ydeg + 1)])
        m = np.concatenate([np.arange(-ll, ll + 1) for ll in range(ydeg + 1)])
        j = m + l
        i = l - m
        q = term[j, i]
        Q = term[j[:, None] + j[None, :], i[:, None] + i[None, :]]
        assert q.shape == (N,) and Q.shape == (N, N)
        _CONST_CACHE[key] = (q, Q)
    return _CONST_CACHE[key]


def wigner_integral_tensors(R, q, Q, ydeg):
    """integrals.py:116-124: U, t[l], T[l]."""
    neig = 2 * ydeg + 1
    U = matrix_sqrt(Q, neig=neig)
    t = [np.dot(R[l], q[l ** 2:(l + 1) ** 2]) for l in range(ydeg + 1)]
    T = [np.swapaxes(np.dot(R[l], U[l ** 2:(l + 1) ** 2]), 1, 2) for l in range(ydeg + 1)]
    return U, t, T


def wigner_first_moment(t, e, ydeg):
    """integrals.py:126-131."""
    mu = np.zeros((ydeg + 1) ** 2)
    for l in range(ydeg + 1):
        i = slice(l ** 2, (l + 1) ** 2)
        mu[i] = np.dot(t[l], e[i])
    return mu


def wigner_second_moment(T, eigE, ydeg):
    """integrals.py:133-151."""
    N = (ydeg + 1) ** 2
    neig = 2 * ydeg + 1
    sqrtC = np.zeros((N, neig, eigE.shape[-1]))
    for l in range(ydeg + 1):
        i = slice(l ** 2, (l + 1) ** 2)
        sqrtC[i] = np.dot(T[l], eigE[i])
    sqrtC = sqrtC.reshape(N, -1)
    if sqrtC.shape[1] > N:
        sqrtC = matrix_sqrt(sqrtC @ sqrtC.T)
    return sqrtC


def longitude_tensors(ydeg):
    key = ("lonT", ydeg)
    if key not in _CONST_CACHE:
        R = wigner_poly_R(ydeg, cos_alpha=1, sin_alpha=0, cos_gamma=1, sin_gamma=0)
        q, Q = longitude_qQ(ydeg)
        _CONST_CACHE[key] = wigner_integral_tensors(R, q, Q, ydeg)
    return _CONST_CACHE[key]


def flux_G(ydeg):
    """flux.py:107-137: G[j_row, i_col] = _G(i_col, j_row)."""
    key = ("G", ydeg)
    if key not in _CONST_CACHE:
        n = 4 * ydeg + 1

        def _G(j, i):
            return 2 * _gamma(1 + 0.5 * i) * _gamma(1 + 0.5 * j) / _gamma(0.5 * (4 + i + j)) - (
                2 ** (1 - 0.5 * i) / (2 + i)
            ) * _sp_hyp2f1(1 + 0.5 * i, -0.5 * j, 2 + 0.5 * i, 0.5)

        _CONST_CACHE[key] = np.array([[_G(i, j) for i in range(n)] for j in range(n)])
    return _CONST_CACHE[key]


def flux_precompute(ydeg):
    """flux.py:123-179: wnp[l], Wnp."""
    key = ("Wnp", ydeg)
    if key not in _CONST_CACHE:
        N = (ydeg + 1) ** 2
        R = wigner_poly_R(ydeg, 0, 1, 0, -1)
        G = flux_G(ydeg)
        wnp = []
        for l in range(ydeg + 1):
            m = np.arange(-l, l + 1)
            wnp.append(R[l] @ G[l - m, l + m])
        Qt = np.empty((2 * ydeg + 1, 2 * ydeg + 1, 2 * ydeg + 1, N))
        for l1 in range(ydeg + 1):
            k = np.arange(l1 ** 2, (l1 + 1) ** 2)
            k0 = np.arange(2 * l1 + 1).reshape(-1, 1)
            for p in range(N):
                l2 = int(np.floor(np.sqrt(p)))
                j = np.arange(l2 ** 2, (l2 + 1) ** 2)
                j0 = np.arange(2 * l2 + 1).reshape(1, -1)
                L = R[l1][l1, k - l1 ** 2] @ G[k0 + j0, 2 * l1 - k0 + 2 * l2 - j0]
                Rr = R[l2][j - l2 ** 2, p - l2 ** 2].T
                Qt[l1, : 2 * l1 + 1, : 2 * l2 + 1, p] = L @ Rr
        Wnp = np.empty((N, N))
        for l1 in range(ydeg + 1):
            i = np.arange(l1 ** 2, (l1 + 1) ** 2)
            for l2 in range(ydeg + 1):
                j = np.arange(l2 ** 2, (l2 + 1) ** 2)
                Wnp[i.reshape(-1, 1), j.reshape(1, -1)] = Qt[l1, : 2 * l1 + 1, l2, j].T
        _CONST_CACHE[key] = (wnp, Wnp)
    return _CONST_CACHE[key]


def alpha_beta_series(z, order=20):
    """ops/norm/norm.py:26-44 (value lanes)."""
    fac = 1.0
    alpha = 0.0
    beta = 0.0
    for n in range(0, order + 1):
        alpha += fac
        beta += 2 * n * fac
        fac *= z * (2 * n + 3)
    return alpha, beta


# ----------------------------------------------------------------------------------------------
# The process
# ----------------------------------------------------------------------------------------------
# temporal.py:8-16
def ExpSquaredKernel(t1, t2, tau):
    dt = np.abs(np.reshape(t1, (-1, 1)) - np.reshape(t2, (1, -1)))
    return np.exp(-(dt ** 2) / (2 * tau))


def Matern32Kernel(t1, t2, tau):
    dt = np.abs(np.reshape(t1, (-1, 1)) - np.reshape(t2, (1, -1)))
    x = np.sqrt(3) * dt / tau
    return (1 + x) * np.exp(-x)


def spot_profile_mean(theta, r, dr, sfac=300):
    """size.py:55-62 (Spot.get_e): the sigmoid spot profile averaged over a uniform radius prior."""
    with np.errstate(over="ignore"):
        chim = np.exp(sfac * (r - dr - theta))
        chip = np.exp(sfac * (r + dr - theta))
        return 1.0 / (2 * dr * sfac) * np.log((1 + chim) / (1 + chip))


def spot_profile_second_moment(theta, r, dr, sfac=300, cutoff=1.5):
    """size.py:64-89 (Spot.get_eigE up to the matrix square root): E[b(theta_i) b(theta_j)] under the
    uniform radius prior, evaluated for theta < cutoff (r + dr) only and zero beyond."""
    kmax = int(np.argmax(theta / (r + dr) > cutoff))
    t = theta[:kmax].reshape(1, -1)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        chim = np.exp(sfac * (r - dr - t))
        chip = np.exp(sfac * (r + dr - t))
        ex = np.exp(sfac * (t - t.T))
        term = np.log(1 + chim) - np.log(1 + chip)
        C0 = (ex * term - term.T) / (1 - ex + 1.0e-15)
        C0[np.arange(kmax), np.arange(kmax)] = (1 / (1 + chip) + chim / (1 + chim) - term - 1).reshape(-1)
    C0 /= 2 * dr * sfac
    C = np.zeros((theta.shape[0], theta.shape[0]))
    C[:kmax, :kmax] = C0
    return C


class OracleProcess(object):
    """Eager NumPy restatement of ``StarryProcess`` (sp.py:38-284) for the lnlike hot path."""

    def __init__(self, r=DEFAULTS["r"], c=DEFAULTS["c"], n=DEFAULTS["n"], mu=None, sigma=None,
                 a=None, b=None, ydeg=15, udeg=2,
                 marginalize_over_inclination=DEFAULTS["marginalize_over_inclination"],
                 normalized=DEFAULTS["normalized"], covpts=DEFAULTS["covpts"], native="port",
                 skip_longitude_eigh=False, q_ulp_noise_seed=None, tau=None,
                 temporal_kernel=Matern32Kernel, dr=None, **kwargs):
        self.nat = get_native(native)
        # sp.py:225-232
        self.tau = None if tau is None else float(check_bounds("tau", tau, 0, np.inf))
        self.temporal_kernel = temporal_kernel
        self.ydeg = int(ydeg)
        self.udeg = int(udeg)
        assert self.ydeg >= 5 or kwargs.pop("allow_low_ydeg", False)
        self.N = (self.ydeg + 1) ** 2
        self.covpts = int(covpts)
        self.normalized = normalized
        self.marg = marginalize_over_inclination
        self.normN = kwargs.get("normalization_order", DEFAULTS["normalization_order"])
        self.normzmax = kwargs.get("normalization_zmax", DEFAULTS["normalization_zmax"])
        self.z = None

        # sp.py:204-222
        if mu is None and sigma is None:
            a = DEFAULTS["a"] if a is None else a
            b = DEFAULTS["b"] if b is None else b
        elif (a is None and b is None) and (mu is not None and sigma is not None):
            a, b = gauss2beta(mu, sigma)
        else:
            raise ValueError("Must provide either `a` and `b` *or* `mu` and `sigma`.")

        ydeg = self.ydeg
        ang = np.pi / 180
        # --- SizeIntegral (delta prior), size.py:93-115
        theta, Bp, idx = spot_Bp(ydeg)
        self.r = float(check_bounds("r", r * ang, 0, 0.5 * np.pi))
        if dr is None:
            zz = 300 * (theta - self.r)           # size.py:45-47 (sfac = 300)
            bprof = 1 / (1 + np.exp(-zz)) - 1
            q_size = np.zeros(self.N)
            q_size[idx] = Bp @ bprof
            eig_size = q_size.reshape(-1, 1)
            self.dr = 0.0
        else:
            # uniform prior on the spot radius over [r - dr, r + dr], size.py:116-125
            self.dr = float(check_bounds("dr", dr * ang, 0, 0.5 * np.pi))
            q_size = np.zeros(self.N)
            q_size[idx] = Bp @ spot_profile_mean(theta, self.r, self.dr)
            self.Etilde = Bp @ spot_profile_second_moment(theta, self.r, self.dr) @ Bp.T
            eig_size = np.zeros((self.N, self.N))
            eig_size[np.ix_(idx, idx)] = matrix_sqrt(self.Etilde)
        self.q_size = q_size

        # --- LatitudeIntegral, latitude.py:171-212
        abmin = kwargs.get("abmin", DEFAULTS["abmin"])
        a = float(check_bounds("a", a, 0, 1))
        a = abmin if a < abmin else a
        b = float(check_bounds("b", b, 0, 1))
        b = abmin if b < abmin else b
        self.a, self.b = a, b
        lam = kwargs.get("log_alpha_max", DEFAULTS["log_alpha_max"])
        lbm = kwargs.get("log_beta_max", DEFAULTS["log_beta_max"])
        self.alpha = np.exp(a * lam)
        self.beta = np.exp(np.log(0.5) + b * (lbm - np.log(0.5)))
        R_lat = wigner_poly_R(ydeg, 0, 1, 0, -1)
        q_lat, Q_lat = self.nat.latitude(ydeg, self.udeg, self.alpha, self.beta)
        if q_ulp_noise_seed is not None:
            # conditioning probe (NOT part of the reference): perturb the latitude moment matrix by
            # symmetric relative noise of one unit roundoff, the size of the error any eigensolver /
            # BLAS / CPU micro-architecture already commits on it.  The spread of lnlike under this
            # probe is the reference's own reproducibility floor for these hyperparameters.
            nrng = np.random.default_rng(q_ulp_noise_seed)
            E = nrng.standard_normal(Q_lat.shape)
            Q_lat = Q_lat * (1.0 + 1.1e-16 * (E + E.T))
        self.q_lat, self.Q_lat = q_lat, Q_lat
        U_lat, t_lat, T_lat = wigner_integral_tensors(R_lat, q_lat, Q_lat, ydeg)
        self.U_lat = U_lat

        # --- LongitudeIntegral, longitude.py:9-49
        U_lon, t_lon, T_lon = longitude_tensors(ydeg)

        # --- ContrastIntegral, contrast.py:9-33
        self.c = float(c)
        self.n = float(check_bounds("n", n, 0, np.inf))
        mom1_lat = wigner_first_moment(t_lat, q_size, ydeg)
        mom1 = wigner_first_moment(t_lon, mom1_lat, ydeg)
        self.sqrtC_lat = wigner_second_moment(T_lat, eig_size, ydeg)
        if skip_longitude_eigh:
            N = self.N
            neig = 2 * ydeg + 1
            sq = np.zeros((N, neig, self.sqrtC_lat.shape[-1]))
            for l in range(ydeg + 1):
                i = slice(l ** 2, (l + 1) ** 2)
                sq[i] = np.dot(T_lon[l], self.sqrtC_lat[i])
            eig_mom2 = sq.reshape(N, -1)
        else:
            eig_mom2 = wigner_second_moment(T_lon, self.sqrtC_lat, ydeg)
        mom2 = eig_mom2 @ eig_mom2.T
        self.mom1 = mom1
        self.mean_ylm = np.pi * self.c * self.n * mom1
        cov = (np.pi * self.c) ** 2 * self.n * (mom2 - np.outer(mom1, mom1))
        lamv = np.ones(self.N) * kwargs.get("epsy", DEFAULTS["epsy"])
        lamv[15 ** 2:] = kwargs.get("epsy15", DEFAULTS["epsy15"])
        self.cov_ylm = cov + np.diag(lamv)
        self._cho = None

        # --- FluxIntegral.__init__, flux.py:55-62
        self._rx90 = self.nat.Rx(ydeg, self.udeg, 0.5 * np.pi)
        self.ez = self._dotRx(self.mean_ylm.reshape(1, -1), self._rx90).T
        mom2y = np.ascontiguousarray(self.cov_ylm + np.outer(self.mean_ylm, self.mean_ylm))
        tmp = np.ascontiguousarray(self._dotRx(mom2y, self._rx90).T)
        self.Ez = self._dotRx(tmp, self._rx90)

    # latitude.py:221-241 (_compute_mu_and_sigma) and 281-316 (_log_jac); sigma_max of latitude.py:194
    def log_jac(self, sigma_max=DEFAULTS["sigma_max"]):
        al, be = self.alpha, self.beta
        term = 4 * al ** 2 - 8 * al - 6 * be + 4 * al * be + be ** 2 + 5
        mu = 2 * np.arctan(np.sqrt(2 * al + be - 2 - np.sqrt(term)))
        term = 1 - al + be + (be - 1) * np.cos(mu) + (al - 1) / np.cos(mu) ** 2
        sigma = np.sqrt(np.sin(mu) ** 2 / term)
        with np.errstate(divide="ignore", invalid="ignore"):
            lj = np.log(np.abs(
                (al * be * (1 + np.cos(mu)) ** 3 * np.sin(2 * mu) ** 3)
                / (sigma * (-3 + 2 * al + be + (-1 + 2 * al + be) * np.cos(mu))
                   * (2 * (-1 + al + be) + 3 * (-1 + be) * np.cos(mu)
                      - 2 * (-1 + al - be) * np.cos(2 * mu) + (-1 + be) * np.cos(3 * mu)) ** 2)))
        return -np.inf if sigma > sigma_max * np.pi / 180 else float(lj)

    # sp.py:265-271
    @property
    def cho_cov_ylm(self):
        if self._cho is None:
            self._cho = cho_factor(self.cov_ylm)
        return self._cho

    # flux.py:74-86
    def _dotRx(self, M, rx):
        f = np.zeros_like(M)
        for l in range(self.ydeg + 1):
            Rxl = rx[nwig(l - 1):nwig(l)].reshape(2 * l + 1, 2 * l + 1)
            f[:, l ** 2:(l + 1) ** 2] = M[:, l ** 2:(l + 1) ** 2] @ Rxl
        return f

    def _rTA1(self, u):
        if self.udeg > 0:
            return self.nat.rTA1L(self.ydeg, self.udeg, np.asarray(u, dtype=float)[: self.udeg])
        return self.nat.rTA1(self.ydeg, self.udeg)

    # flux.py:278-281, 88-105
    def design_matrix(self, t, i=DEFAULTS["i"], p=DEFAULTS["p"], u=(0.0, 0.0)):
        t = np.asarray(t, dtype=float).reshape(-1)
        inc = float(check_bounds("i", i * np.pi / 180, 0, 0.5 * np.pi))
        p = float(check_bounds("p", p, 0, np.inf))
        theta = 2 * np.pi * np.mod(t / p, 1.0)
        M = np.tile(self._rTA1(u), (theta.shape[0], 1))
        M = self._dotRx(M, self.nat.Rx(self.ydeg, self.udeg, -inc))
        M = self.nat.tensordotRz(self.ydeg, self.udeg, M, theta)
        M = self._dotRx(M, self._rx90)
        return M

    # flux.py:181-231, 283-343
    def _flux_mean_cov(self, t, i, p, u):
        t = np.asarray(t, dtype=float).reshape(-1)
        ydeg = self.ydeg
        if self.marg:
            check_bounds("i", i * np.pi / 180, 0, 0.5 * np.pi)
            p = float(check_bounds("p", p, 0, np.inf))
            rTA1 = self._rTA1(u)
            wnp, Wnp = flux_precompute(ydeg)
            w = [rTA1[l ** 2:(l + 1) ** 2] @ wnp[l] for l in range(ydeg + 1)]
            m0 = np.array([l ** 2 + l for l in range(ydeg + 1)])
            Z = np.outer(rTA1[m0], rTA1[m0])
            W = np.zeros((self.N, self.N))
            for l1 in range(ydeg + 1):
                for l2 in range(ydeg + 1):
                    W[l1 ** 2:(l1 + 1) ** 2, l2 ** 2:(l2 + 1) ** 2] = (
                        Wnp[l1 ** 2:(l1 + 1) ** 2, l2 ** 2:(l2 + 1) ** 2] * Z[l1, l2])
            mean = np.sum([np.dot(w[l], self.ez[l ** 2:(l + 1) ** 2]) for l in range(ydeg + 1)])
            var = (np.tensordot(W, self.Ez) - mean ** 2) * np.eye(1)
            dx = 2 * np.pi / self.covpts
            xp = np.arange(-dx, 2 * np.pi + 2.5 * dx, dx)
            mom2 = self.nat.special_tensordotRz(ydeg, self.udeg, W, self.Ez, xp)
            yp = mom2 - mean ** 2
            y0, y1, y2, y3 = yp[:-3], yp[1:-2], yp[2:-1], yp[3:]
            a0 = y1
            a1 = -y0 / 3.0 - 0.5 * y1 + y2 - y3 / 6.0
            a2 = 0.5 * (y0 + y2) - y1
            a3 = 0.5 * ((y1 - y2) + (y3 - y0) / 3.0)
            self._kernel_grid = (xp, yp)
            self._interp = (dx, xp, a0, a1, a2, a3, p)
            # flux.py:256-276
            theta = 2 * np.pi * np.mod(t / p, 1.0)
            x = np.abs(theta[:, None] - theta[None, :]).reshape(-1)
            inds = np.floor(x / dx).astype("int64")
            x0 = (x - xp[inds + 1]) / dx
            cov = (a0[inds] + a1[inds] * x0 + a2[inds] * x0 ** 2
                   + a3[inds] * x0 ** 3).reshape(theta.shape[0], theta.shape[0])
            if theta.shape[0] == 1:
                cov = var
            return float(mean), cov
        A = self.design_matrix(t, i, p, u)
        mean = float(np.dot(A, self.mean_ylm)[0])
        cov = np.dot(np.dot(A, self.cov_ylm), A.T)
        return mean, cov

    # sp.py:643-672
    def mean(self, t, i=DEFAULTS["i"], p=DEFAULTS["p"], u=(0.0, 0.0)):
        t = np.asarray(t, dtype=float).reshape(-1)
        if self.normalized:
            return np.zeros_like(t)
        m, _ = self._flux_mean_cov(t, i, p, u)
        return m * np.ones_like(t)

    # sp.py:674-727
    def cov(self, t, i=DEFAULTS["i"], p=DEFAULTS["p"], u=(0.0, 0.0)):
        mean, cov = self._flux_mean_cov(t, i, p, u)
        if self.tau is not None:   # sp.py:697-698
            cov = cov * self.temporal_kernel(t, t, self.tau)
        if self.normalized:
            return self._normalize(1.0 + mean, cov)
        return cov

    def _normalize(self, mu, Sig):
        K = Sig.shape[0]
        j = np.ones((K, 1))
        m = np.mean(Sig)
        q = np.dot(Sig, j) / (K * m)
        self.z = m / mu ** 2
        p = j - q
        alpha, beta = alpha_beta_series(self.z, self.normN)
        ppT = np.dot(p, p.T)
        qqT = np.dot(q, q.T)
        return (alpha / mu ** 2) * Sig + self.z * ((alpha + beta) * ppT - alpha * qqT)

    # sp.py:1052-1188
    def log_likelihood(self, t, flux, data_cov, i=DEFAULTS["i"], p=DEFAULTS["p"], u=(0.0, 0.0),
                       baseline_mean=0.0, baseline_var=0.0, return_parts=False):
        gp_mean = self.mean(t, i=i, p=p, u=u)
        gp_cov = np.array(self.cov(t, i=i, p=p, u=u))
        K = gp_mean.shape[0]
        data_cov = np.asarray(data_cov, dtype=float)
        if data_cov.ndim == 0:
            C = data_cov * np.eye(K)
        elif data_cov.ndim == 1:
            C = np.diag(data_cov)
        else:
            C = data_cov
        gp_cov = gp_cov + C
        gp_cov = gp_cov + baseline_var
        L = cho_factor(gp_cov)
        mean = np.reshape(gp_mean + baseline_mean, (K, 1))
        r = np.reshape(np.transpose(np.asarray(flux, dtype=float)), (K, -1)) - mean
        M = r.shape[1]
        lnlike = -0.5 * np.sum(r * cho_solve(L, r))
        with np.errstate(invalid="ignore"):
            lnlike -= M * np.sum(np.log(np.diag(L)))
        lnlike -= 0.5 * K * M * np.log(2 * np.pi)
        if self.normalized and self.z > self.normzmax:
            lnlike = -np.inf
        if np.isnan(lnlike):
            lnlike = -np.inf
        if return_parts:
            return float(lnlike), gp_cov, L
        return float(lnlike)

    # sp.py:489-509 with an explicit standard-normal matrix `u` of shape (N, nsamples)
    def sample_ylm(self, unit_normals):
        return np.transpose(self.mean_ylm[:, None] + np.dot(self.cho_cov_ylm, unit_normals))

    # ---------------------------------------------------------------- SURVEY 8(f) rank 2
    # sp.py:729-765 with the standard-normal matrix `unit_normals` (nt, nsamples) given
    def sample(self, t, unit_normals, i=DEFAULTS["i"], p=DEFAULTS["p"], u=(0.0, 0.0), eps=1e-8):
        t = np.asarray(t, dtype=float).reshape(-1)
        cho_cov = cho_factor(self.cov(t, i, p, u) + eps * np.eye(t.shape[0]))
        return np.transpose(self.mean(t, i, p, u)[:, None] + np.dot(cho_cov, unit_normals))

    # sp.py:767-922
    def predict(self, t, flux, data_cov, t_sample=None, i=DEFAULTS["i"], p=DEFAULTS["p"],
                u=(0.0, 0.0), baseline_mean=0.0, baseline_var=0.0):
        if self.normalized:
            raise NotImplementedError("Method not implemented when the flux is normalized.")
        t = np.asarray(t, dtype=float).reshape(-1)
        cov_t = self.cov(t, i, p, u)
        if t_sample is None:
            ts, cov_ts = t, cov_t
        else:
            ts = np.asarray(t_sample, dtype=float).reshape(-1)
            cov_ts = self.cov(ts, i, p, u)
        y = np.asarray(flux, dtype=float) - baseline_mean
        data_cov = np.asarray(data_cov, dtype=float)
        if data_cov.ndim == 0:
            data_cov = data_cov * np.eye(t.shape[0])
        elif data_cov.ndim == 1:
            data_cov = np.diag(data_cov)
        mean, _ = self._flux_mean_cov(np.array([0.0]), i, p, u)
        K_t_t = cov_t + data_cov + baseline_var
        K_ts_ts = cov_ts + baseline_var
        if self.marg:
            self._flux_mean_cov(t, i, p, u)
            dx, xp, a0, a1, a2, a3, per = self._interp
            theta_t = 2 * np.pi * np.mod(t / per, 1.0)
            theta_ts = 2 * np.pi * np.mod(ts / per, 1.0)
            x = np.abs(theta_ts[:, None] - theta_t[None, :]).reshape(-1)
            inds = np.floor(x / dx).astype("int64")
            x0 = (x - xp[inds + 1]) / dx
            K_ts_t = (a0[inds] + a1[inds] * x0 + a2[inds] * x0 ** 2
                      + a3[inds] * x0 ** 3).reshape(theta_ts.shape[0], theta_t.shape[0])
        else:
            A_ts = self.design_matrix(ts, i, p, u)
            A_t = self.design_matrix(t, i, p, u)
            K_ts_t = np.dot(np.dot(A_ts, self.cov_ylm), A_t.T)
        if self.tau is not None:   # sp.py:893-894
            K_ts_t = K_ts_t * self.temporal_kernel(ts, t, self.tau)
        K_ts_t = K_ts_t + baseline_var
        cho_K = cho_factor(K_t_t)
        mu = mean + np.dot(K_ts_t, cho_solve(cho_K, y - mean))
        K = K_ts_ts - np.dot(K_ts_t, cho_solve(cho_K, K_ts_t.T))
        return mu, K

    # sp.py:924-1002 with `ts = t_sample` (the reference body uses an undefined name `ts`) and the
    # standard-normal matrix `unit_normals` (nts, nsamples) given
    def sample_conditional(self, t, flux, data_cov, unit_normals, t_sample=None, eps=1e-8, **kw):
        mu, K = self.predict(t, flux, data_cov, t_sample=t_sample, **kw)
        cho_K = cho_factor(K + eps * np.eye(K.shape[0]))
        return np.transpose(mu[:, None] + np.dot(cho_K, unit_normals))

    # sp.py:518-641 with the standard-normal matrix `unit_normals` (N, nsamples) given
    def sample_ylm_conditional(self, t, flux, data_cov, unit_normals, i=DEFAULTS["i"],
                               p=DEFAULTS["p"], u=(0.0, 0.0), baseline_mean=0.0, baseline_var=0.0):
        if self.normalized:
            raise NotImplementedError("Method not implemented when the flux is normalized.")
        if self.tau is not None:
            raise NotImplementedError("Method not implemented for time-variable maps.")
        flux = np.asarray(flux, dtype=float)
        data_cov = np.asarray(data_cov, dtype=float)
        if data_cov.ndim == 0:
            C = data_cov * np.eye(flux.shape[0])
        elif data_cov.ndim == 1:
            C = np.diag(data_cov)
        else:
            C = np.array(data_cov)
        C = C + baseline_var
        cho_C = cho_factor(C)
        A = self.design_matrix(t, i, p, u)
        CInvA = cho_solve(cho_C, A)
        LInv = cho_solve(self.cho_cov_ylm, np.eye(self.N))      # sp.py:267-270
        LInvmu = cho_solve(self.cho_cov_ylm, self.mean_ylm)     # sp.py:271
        W = np.dot(A.T, CInvA) + LInv
        cho_W = cho_factor(W)
        M = cho_solve(cho_W, CInvA.T)
        ymu = np.dot(M, flux - baseline_mean) + cho_solve(cho_W, LInvmu)
        ycov = cho_solve(cho_W, np.eye(self.N))
        cho_ycov = cho_factor(ycov)
        return np.transpose(ymu[:, None] + np.dot(cho_ycov, unit_normals))

    # sp.py:510-516 + ops/sample.py:24-33 (SampleYlmTemporalOp: y_i = L_t U_i L_y^T; the reference
    # does NOT add mean_ylm on this branch), with U (nsamples, nt, N) given
    def sample_ylm_temporal(self, t, unit_normals):
        cov_t = self.temporal_kernel(t, t, self.tau)
        cho_cov_t = cho_factor(cov_t)
        Ly = self.cho_cov_ylm
        return np.einsum("km,imj,nj->ikn", cho_cov_t, unit_normals, Ly)

    # sp.py:1237-1283
    def flux(self, y, t, i=DEFAULTS["i"], p=DEFAULTS["p"], u=(0.0, 0.0)):
        y = np.asarray(y, dtype=float)
        A = self.design_matrix(t, i, p, u)
        F = np.tensordot(A, y, axes=[[1], [y.ndim - 1]])
        if self.tau is not None:
            flux = np.diagonal(F, axis1=0, axis2=y.ndim - 1)
        else:
            flux = np.transpose(F)
        if self.normalized:
            flux = (1.0 + flux) / np.reshape(np.mean(1.0 + flux, axis=-1), (-1, 1)) - 1.0
        return flux

    # sp.py:1190-1198
    def __add__(self, other):
        return OracleProcessSum(self, other)


class OracleProcessSum(OracleProcess):
    """sp.py:1335-1400."""

    def __init__(self, first, second):
        assert first.normalized == second.normalized and first.marg == second.marg
        assert first.covpts == second.covpts and first.tau is None and second.tau is None
        self.__dict__.update(first.__dict__)
        self.mean_ylm = first.mean_ylm + second.mean_ylm
        self.cov_ylm = first.cov_ylm + second.cov_ylm
        self._cho = None
        # FluxIntegral.__init__ (flux.py:55-62) on the summed moments
        self.ez = self._dotRx(self.mean_ylm.reshape(1, -1), self._rx90).T
        mom2y = np.ascontiguousarray(self.cov_ylm + np.outer(self.mean_ylm, self.mean_ylm))
        tmp = np.ascontiguousarray(self._dotRx(mom2y, self._rx90).T)
        self.Ez = self._dotRx(tmp, self._rx90)
