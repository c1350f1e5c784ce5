"""
TEST INFRASTRUCTURE ONLY (see oracle/sp_oracle.py).  Gradient of the log-likelihood with respect to
the hyperparameters (r, a, b, c, n), SURVEY.md section 8(f) rank 4.

What the reference does: Theano reverse mode through its ops --
``computeLatitudeIntegrals`` derivative lanes (ops/include/latitude.h:22-173), ``eigh_grad``
(ops/include/eigh.h:19-65), ``computeTensordotRzGradient`` / ``computeSpecialTensordotRzGradient``
(ops/include/wigner.h:345-404, 465-531), ``Solve.L_op`` / ``Cholesky`` (math.py:40-72).  Theano is not
installable here, so this module restates the same derivative analytically in NumPy and is PINNED
the way the reference pins its own gradient: against finite differences of the forward function
(``theano.gradient.verify_grad`` in tests/test_lnlike.py:105-136, tolerance 1e-4) -- here central
differences of the UNMODIFIED reference's ``log_likelihood`` run through oracle/theano_stub
(tests/test_oracle_grad.py), and of this oracle's own forward.

Structure (the same one the CUDA path uses):
  * the only non-linear, hyperparameter-dependent inputs of the Ylm moments are the Beta moments of
    the latitude distribution (-> Q_lat, q_lat; derivative lanes as latitude.h:48-60, 112-143), the
    spot profile (-> e, E) and the scalars c, n;
  * everything from those to (mean_ylm, cov_ylm) is linear / bilinear, and the matrix square roots
    of integrals.py:116-151 are only a compression: C_lat = E o (P Q_lat P^T), C_lon = sum_e T_e C_lat
    T_e^T, so the tangents need no eigen-derivative;
  * (mean_ylm, cov_ylm) -> (gp_mean, K) is linear in cov_ylm and quadratic in mean_ylm, followed by
    the normalisation series (sp.py:705-727), whose tangent is analytic;
  * d lnlike = 1/2 alpha^T dK alpha - 1/2 tr(K^-1 dK) + (alpha^T 1) d(mean),  alpha = K^-1 r.
"""
import numpy as np

from . import sp_oracle as so

PARAMS = ("r", "a", "b", "c", "n")


# ----------------------------------------------------------------------------------------------
# latitude integrals with derivative lanes (latitude.h:48-60, 112-172), even/even terms only: the
# scatter of latitude.h:146-172 never reads an odd second index, and an even second index implies an
# even first one, so the hypergeometric "F" lane (latitude.h:63-109) does not enter.
# ----------------------------------------------------------------------------------------------
def latitude_with_grad(ydeg, alpha, beta):
    n = 4 * ydeg + 1
    B = np.zeros(n)
    dBa = np.zeros(n)
    dBb = np.zeros(n)
    B[0] = 1.0
    for k in range(1, n):
        c1 = 1.0 / (alpha + beta + k - 1.0)
        c2 = (alpha + k - 1.0) * c1
        c3 = beta * c1 * c1
        c4 = (1 - k - alpha) * c1 * c1
        B[k] = c2 * B[k - 1]
        dBa[k] = c3 * B[k - 1] + c2 * dBa[k - 1]
        dBb[k] = c4 * B[k - 1] + c2 * dBb[k - 1]
    nh = 2 * ydeg + 1
    term = np.zeros((3, n, n))
    for i2 in range(nh):
        for j2 in range(nh):
            if i2 + j2 > 2 * ydeg:
                continue
            fac1 = 1.0
            acc = np.zeros(3)
            for k1 in range(i2 + 1):
                fac2 = fac1
                for k2 in range(j2 + 1):
                    acc[0] += fac2 * B[k1 + k2]
                    acc[1] += fac2 * dBa[k1 + k2]
                    acc[2] += fac2 * dBb[k1 + k2]
                    fac2 *= (k2 - j2) / (k2 + 1.0)
                fac1 *= (i2 - k1) / (k1 + 1.0)
            term[:, 2 * i2, 2 * j2] = acc
    l = np.concatenate([np.full(2 * ll + 1, ll) for ll in range(ydeg + 1)])
    m = np.concatenate([np.arange(-ll, ll + 1) for ll in range(ydeg + 1)])
    j, i = m + l, l - m
    q = term[:, j, i] * (0.5 ** l)[None, :]
    Q = term[:, j[:, None] + j[None, :], i[:, None] + i[None, :]] * (0.5 ** (l[:, None] + l[None, :]))[None]
    return q, Q     # q[0], Q[0]: values; [1]: d/dalpha; [2]: d/dbeta


# ----------------------------------------------------------------------------------------------
# the linear mid-section: (e, E, q_lat, Q_lat) -> (mom1, second moment C_lon)
# ----------------------------------------------------------------------------------------------
class LinearMoments(object):
    def __init__(self, ydeg=15):
        self.ydeg = ydeg
        self.N = (ydeg + 1) ** 2
        self.R_lat = so.wigner_poly_R(ydeg, 0, 1, 0, -1)
        self.R_lon = so.wigner_poly_R(ydeg, cos_alpha=1, sin_alpha=0, cos_gamma=1, sin_gamma=0)
        self.U_lon, self.t_lon, self.T_lon = so.longitude_tensors(ydeg)
        self.l_of = np.concatenate([np.full(2 * ll + 1, ll) for ll in range(ydeg + 1)])
        # P = blockdiag_l R_lat[l][:, m = 0, :]   (rows m', columns k): C_lat = E o (P Q_lat P^T).
        # P carries the polynomial Wigner coefficients (up to 7.8e7 at l = 15) and Q_lat has exact
        # rank 2 ydeg + 1: evaluated naively in fp64 the product cancels catastrophically (1e-3 of
        # the largest entry).  As in the product's constant tables, Q_lat is therefore restricted to
        # its exact range first -- Z = orthonormal basis of the monomials c^j s^(2l-j) promoted to
        # degree 2 ydeg -- and the folded operator H = P Z is formed in extended precision:
        #     C_lat = E o (H S H^T),   S = Z^T Q_lat Z   ((2 ydeg + 1)^2, well conditioned).
        import math

        neig0 = 2 * ydeg + 1
        V = np.zeros((self.N, neig0))
        for l in range(ydeg + 1):
            for k in range(2 * l + 1):
                for tt_ in range(ydeg - l + 1):
                    V[l * l + k, k + 2 * tt_] = math.comb(ydeg - l, tt_)
        self.Z = np.linalg.svd(V, full_matrices=False)[0]
        H = np.zeros((self.N, neig0))
        for l in range(ydeg + 1):
            s = slice(l * l, (l + 1) ** 2)
            H[s] = (self.R_lat[l][:, l, :].astype(np.longdouble)
                    @ self.Z[s].astype(np.longdouble)).astype(np.float64)
        self.H = H
        # T_e = blockdiag_l T_lon[l][:, e, :]
        neig = 2 * ydeg + 1
        self.Te = np.zeros((neig, self.N, self.N))
        for l in range(ydeg + 1):
            s = slice(l * l, (l + 1) ** 2)
            self.Te[:, s, s] = np.swapaxes(self.T_lon[l], 0, 1)

    def first_moment(self, e_l, q_lat):
        """integrals.py:126-131 twice (latitude then longitude); e_l: (ydeg+1,) size moments."""
        ydeg = self.ydeg
        m1 = np.zeros(self.N)
        for l in range(ydeg + 1):
            s = slice(l * l, (l + 1) ** 2)
            t_lat = np.dot(self.R_lat[l], q_lat[s])          # (m', m)
            m1_lat = t_lat[:, l] * e_l[l]                      # e is non-zero at m = 0 only
            m1[s] = np.dot(self.t_lon[l], m1_lat)
        return m1

    def second_moment(self, E, Q_lat):
        """C_lon = sum_e T_e (E[l1, l2] o P Q_lat P^T) T_e^T -- what integrals.py:133-151 computes
        through matrix square roots (sqrtC sqrtC^T), without them (P Q P^T = H S H^T, see above)."""
        S = self.Z.T @ Q_lat @ self.Z
        C_lat = E[np.ix_(self.l_of, self.l_of)] * (self.H @ (0.5 * (S + S.T)) @ self.H.T)
        C = np.zeros((self.N, self.N))
        for e in range(self.Te.shape[0]):
            C += self.Te[e] @ C_lat @ self.Te[e].T
        return C


_LIN = {}


def linear_moments(ydeg=15):
    if ydeg not in _LIN:
        _LIN[ydeg] = LinearMoments(ydeg)
    return _LIN[ydeg]


def moment_tangents(r, a, b, c, n, ydeg=15, abmin=1e-12, lam=10, lbm=10, epsy=1e-12, epsy15=1e-9):
    """(mean_ylm, cov_ylm) and their derivatives with respect to (r [deg], a, b, c, n): returns
    ``mean (N,), cov (N, N), dmean {p: (N,)}, dcov {p: (N, N)}`` (delta prior on the spot radius)."""
    lin = linear_moments(ydeg)
    N = lin.N
    ang = np.pi / 180
    theta, Bp, idx = so.spot_Bp(ydeg)
    rr = r * ang
    z = 300 * (theta - rr)
    sig = 1 / (1 + np.exp(-z))
    e_l = Bp @ (sig - 1)
    de_l = Bp @ (-300.0 * sig * (1 - sig)) * ang          # d e / d r[deg]
    a_ = max(a, abmin)
    b_ = max(b, abmin)
    alpha = np.exp(a_ * lam)
    beta = np.exp(np.log(0.5) + b_ * (lbm - np.log(0.5)))
    q3, Q3 = latitude_with_grad(ydeg, alpha, beta)
    dal = lam * alpha if a > abmin else 0.0               # clamp: derivative vanishes below abmin
    dbe = (lbm - np.log(0.5)) * beta if b > abmin else 0.0
    E = np.outer(e_l, e_l)
    m1 = lin.first_moment(e_l, q3[0])
    C = lin.second_moment(E, Q3[0])
    d_m1 = {"r": lin.first_moment(de_l, q3[0]),
            "a": lin.first_moment(e_l, q3[1]) * dal,
            "b": lin.first_moment(e_l, q3[2]) * dbe}
    d_C = {"r": lin.second_moment(np.outer(de_l, e_l) + np.outer(e_l, de_l), Q3[0]),
           "a": lin.second_moment(E, Q3[1]) * dal,
           "b": lin.second_moment(E, Q3[2]) * dbe}
    sc = (np.pi * c) ** 2 * n
    lamv = np.ones(N) * epsy
    lamv[15 ** 2:] = epsy15
    mean = np.pi * c * n * m1
    cov = sc * (C - np.outer(m1, m1)) + np.diag(lamv)
    dmean, dcov = {}, {}
    for p in ("r", "a", "b"):
        dmean[p] = np.pi * c * n * d_m1[p]
        dcov[p] = sc * (d_C[p] - np.outer(d_m1[p], m1) - np.outer(m1, d_m1[p]))
    dmean["c"] = mean / c
    dcov["c"] = 2 * (cov - np.diag(lamv)) / c
    dmean["n"] = mean / n
    dcov["n"] = (cov - np.diag(lamv)) / n
    return mean, cov, dmean, dcov


# ----------------------------------------------------------------------------------------------
# flux side
# ----------------------------------------------------------------------------------------------
def _alpha_beta_with_grad(z, order=20):
    """ops/norm/norm.py:26-44 with the z-derivatives of both series."""
    fac, dfac = 1.0, 0.0
    al = be = dal = dbe = 0.0
    for k in range(order + 1):
        al += fac
        dal += dfac
        be += 2 * k * fac
        dbe += 2 * k * dfac
        dfac = (2 * k + 3) * (fac + z * dfac)
        fac *= z * (2 * k + 3)
    return al, be, dal, dbe


def _flux_from_moments(o, mean_ylm, cov_ylm, t, i, p, u):
    """(gp_mean, K0) of an OracleProcess-like object for GIVEN Ylm moments (flux.py:55-62, 283-343)."""
    o.mean_ylm, o.cov_ylm = mean_ylm, cov_ylm
    o.ez = o._dotRx(mean_ylm.reshape(1, -1), o._rx90).T
    mom2y = np.ascontiguousarray(cov_ylm + np.outer(mean_ylm, mean_ylm))
    tmp = np.ascontiguousarray(o._dotRx(mom2y, o._rx90).T)
    o.Ez = o._dotRx(tmp, o._rx90)
    return o._flux_mean_cov(t, i, p, u)


def lnlike_and_grad(hp, t, flux, data_cov, i=60.0, p=1.0, u=(0.0, 0.0), baseline_mean=0.0,
                    baseline_var=0.0, marginalize_over_inclination=True, normalized=True,
                    native="port"):
    """log-likelihood (sp.py:1052-1188) and its gradient with respect to r [deg], a, b, c, n.
    ``hp``: dict with r, a, b (or mu, sigma), c, n.  Returns ``(lnlike, {p: d lnlike / d p})``."""
    hp = dict(hp)
    if "mu" in hp:
        hp["a"], hp["b"] = so.gauss2beta(hp.pop("mu"), hp.pop("sigma"))
    o = so.OracleProcess(marginalize_over_inclination=marginalize_over_inclination,
                         normalized=normalized, native=native, **hp)
    mean_y, cov_y, dmean_y, dcov_y = moment_tangents(hp["r"], hp["a"], hp["b"], hp["c"], hp["n"])
    t = np.asarray(t, dtype=float).reshape(-1)
    K = t.shape[0]
    one = np.ones(K)

    def flux_side(my, cy):
        return _flux_from_moments(o, my, cy, t, i, p, u)

    gm, K0 = flux_side(mean_y, cov_y)
    # tangents of (gp_mean, K0): the map is linear in cov_ylm and quadratic in mean_ylm, so the
    # central difference along (dmean, dcov) is EXACT (its even part cancels, no third-order term)
    dgm, dK0 = {}, {}
    for q in PARAMS:
        h = 1e-3 * np.abs(cov_y).max() / max(np.abs(dcov_y[q]).max(), 1e-300)
        gp_, Kp_ = flux_side(mean_y + h * dmean_y[q], cov_y + h * dcov_y[q])
        gm_, Km_ = flux_side(mean_y - h * dmean_y[q], cov_y - h * dcov_y[q])
        dgm[q] = (gp_ - gm_) / (2 * h)
        dK0[q] = (Kp_ - Km_) / (2 * h)
    flux_side(mean_y, cov_y)   # restore
    # normalisation (sp.py:705-727) and its tangent
    if normalized:
        mu = 1.0 + gm
        mbar = np.mean(K0)
        qv = K0 @ one / (K * mbar)
        pv = one - qv
        z = mbar / mu ** 2
        al, be, dal, dbe = _alpha_beta_with_grad(z, o.normN)
        Kt = (al / mu ** 2) * K0 + z * ((al + be) * np.outer(pv, pv) - al * np.outer(qv, qv))
        dKt = {}
        for q in PARAMS:
            dmu = dgm[q]
            dmb = np.mean(dK0[q])
            dqv = dK0[q] @ one / (K * mbar) - qv * dmb / mbar
            dz = dmb / mu ** 2 - 2 * mbar * dmu / mu ** 3
            d_al, d_be = dal * dz, dbe * dz
            dKt[q] = ((d_al / mu ** 2 - 2 * al * dmu / mu ** 3) * K0 + (al / mu ** 2) * dK0[q]
                      + dz * ((al + be) * np.outer(pv, pv) - al * np.outer(qv, qv))
                      + z * ((d_al + d_be) * np.outer(pv, pv)
                             - (al + be) * (np.outer(dqv, pv) + np.outer(pv, dqv))
                             - d_al * np.outer(qv, qv)
                             - al * (np.outer(dqv, qv) + np.outer(qv, dqv))))
        mean_vec, dmean_flux = np.zeros(K), {q: 0.0 for q in PARAMS}   # normalised: mean == 0
        zval = z
    else:
        Kt, dKt = K0, dK0
        mean_vec, dmean_flux = gm * one, dgm
        zval = None
    dc = np.asarray(data_cov, dtype=float)
    Cn = dc * np.eye(K) if dc.ndim == 0 else (np.diag(dc) if dc.ndim == 1 else dc)
    Kf = Kt + Cn + baseline_var
    f = np.asarray(flux, dtype=float)
    R = np.reshape(np.transpose(f), (K, -1)) - (mean_vec + baseline_mean).reshape(K, 1)
    M = R.shape[1]
    L = so.cho_factor(Kf)
    Al = so.cho_solve(L, R)
    ll = -0.5 * np.sum(R * Al) - M * np.sum(np.log(np.diag(L))) - 0.5 * K * M * np.log(2 * np.pi)
    Kinv = so.cho_solve(L, np.eye(K))
    Kbar = 0.5 * (Al @ Al.T - M * Kinv)
    grad = {}
    for q in PARAMS:
        grad[q] = float(np.sum(Kbar * dKt[q]) + np.sum(Al) * dmean_flux[q])
    if (normalized and zval > o.normzmax) or np.isnan(ll):
        ll = -np.inf
        grad = {q: 0.0 for q in PARAMS}
    return float(ll), grad
