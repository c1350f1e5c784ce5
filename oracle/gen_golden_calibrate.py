"""
TEST INFRASTRUCTURE ONLY.  Golden values for the batched callers:

    make -C oracle ref && python -m oracle.gen_golden_calibrate

``get_log_prob`` (calibrate/log_prob.py:7-106) builds a *symbolic* Theano function, which the eager
stub cannot trace; its 40 lines of glue (log_prob.py:48-90) are therefore restated here on top of
the UNMODIFIED reference ``StarryProcess`` (mean / cov / log_jac evaluated through
oracle/theano_stub) and the reference's own ``cho_factor`` / ``cho_solve`` (math.py:75-100).
Also emits the inclination grid of calibrate/inclination.py:63-74 for a small case.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import theano_stub  # noqa: E402

U_LD = [0.4, 0.26]


def ev(x):
    return np.array(x.eval() if hasattr(x, "eval") else x, dtype=np.float64)


def ref_log_prob(sp, t, flux, ferr, p, r, a, b, c, n, m, v, i, apply_jac, normalized, marg, u):
    """log_prob.py:37-90 with concrete numbers."""
    from starry_process.math import cho_factor, cho_solve

    K = len(t)
    g = sp.StarryProcess(ydeg=15, r=r, a=a, b=b, c=c, n=n, normalized=normalized,
                         marginalize_over_inclination=marg, covpts=K - 1)
    flux = np.atleast_2d(flux)
    nlc = flux.shape[0]
    gp_mean = ev(g.mean(t, p=p, i=i, u=u))
    gp_cov = ev(g.cov(t, p=p, i=i, u=u))
    R = flux.T - gp_mean.reshape(-1, 1)
    R = R - m
    gp_cov = gp_cov + ferr ** 2 * np.eye(K)
    gp_cov = gp_cov + 10 ** v
    cho = ev(cho_factor(gp_cov))
    CInvR = ev(cho_solve(cho, R))
    ll = -0.5 * np.sum(R * CInvR)
    ll -= nlc * np.sum(np.log(np.diag(cho)))
    ll -= 0.5 * nlc * K * np.log(2 * np.pi)
    if np.isnan(ll):
        ll = -np.inf
    lj = float(ev(g.log_jac()))
    return (ll + lj if apply_jac else ll), lj


def main():
    sp = theano_stub.import_reference()
    gfid = np.load(os.path.join(ROOT, "tests", "golden", "fiducial_nt1000.npz"))
    t = gfid["t"][::4].copy()                 # 250 points, covpts = 249
    fl = gfid["flux_ens_norm"][:3, ::4].copy()
    hp, _, _, _ = bench.synthetic_inputs(4096, seed=1234)
    ns = 12
    a = np.zeros(ns)
    b = np.zeros(ns)
    for s in range(ns):
        a[s], b[s] = sp.gauss2beta(hp["mu"][s], hp["sigma"][s])
    res = dict(t=t, flux=fl, r=hp["r"][:ns], a=a, b=b, c=hp["c"][:ns], n=hp["n"][:ns],
               m=1e-4 * np.arange(ns), v=-3.0 - 0.1 * np.arange(ns), ferr=1e-3, p=1.0, u=np.array(U_LD))
    lp_m = np.zeros(ns)
    lp_c = np.zeros(ns)
    lj = np.zeros(ns)
    for s in range(ns):
        # marginalised over inclination, 3 light curves jointly, fixed baseline (defaults), + jacobian
        lp_m[s], lj[s] = ref_log_prob(sp, t, fl, 1e-3, 1.0, res["r"][s], a[s], b[s], res["c"][s],
                                      res["n"][s], 0.0, 0.0, 60.0, True, True, True, U_LD)
        # conditional, free flux / baseline mean / baseline log-variance / inclination
        lp_c[s], _ = ref_log_prob(sp, t, fl[:1], 1e-3, 1.0, res["r"][s], a[s], b[s], res["c"][s],
                                  res["n"][s], res["m"][s], res["v"][s], 20.0 + 5.0 * s, True, True,
                                  False, U_LD)
        print(s, lp_m[s], lp_c[s], lj[s], flush=True)
    res.update(log_prob_marg=lp_m, log_prob_cond=lp_c, log_jac=lj, inc_cond=20.0 + 5.0 * np.arange(ns))
    # inclination grid: 2 light curves x 2 samples x 5 inclinations (inclination.py:63-74)
    inc = np.array([0.0, 22.5, 45.0, 67.5, 90.0])
    grid = np.zeros((2, 2, inc.size))
    for l in range(2):
        for s in range(2):
            for k, ik in enumerate(inc):
                grid[l, s, k], _ = ref_log_prob(sp, t, fl[l:l + 1], 1e-3, 1.0, res["r"][s], a[s], b[s],
                                                res["c"][s], res["n"][s], 0.0, 0.0, ik, True, True,
                                                False, [0.0, 0.0])
    res.update(inc_grid=inc, lp_grid=grid)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "calibrate_log_prob.npz"), **res)
    print("done")


if __name__ == "__main__":
    main()
