"""
TEST INFRASTRUCTURE ONLY.  Golden fixture for the uniform spot-size prior (``dr``; size.py:55-89,
116-125), produced by the UNMODIFIED reference through oracle/theano_stub:

    make -C oracle ref && python -m oracle.gen_golden_dr

``size_dr.npz``: the reference's own test point (tests/test_size.py: r = 15, dr = 5, ydeg = 15) and a
24-draw sweep (r, dr, mu, sigma, c, n): mean_ylm, diagonal / one row / Frobenius norm of cov_ylm, and
the log-likelihood of the fiducial nt = 300 light curve in all four marginalise x normalise modes.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import theano_stub  # noqa: E402

U_LD = [0.4, 0.26]


def ev(x):
    return np.array(x.eval() if hasattr(x, "eval") else x, dtype=np.float64)


def main():
    SP = theano_stub.import_reference().StarryProcess
    fid = np.load(os.path.join(ROOT, "tests", "golden", "fiducial_nt1000.npz"))
    t = fid["t"][:300].copy()
    f, fn = fid["flux"][:300].copy(), fid["flux_norm"][:300].copy()
    ns = 24
    rng = np.random.default_rng(21)
    hp = dict(r=rng.uniform(10, 30, ns), dr=rng.uniform(1, 9, ns), c=rng.uniform(0.01, 0.15, ns),
              n=rng.uniform(1, 12, ns), mu=rng.uniform(0, 85, ns), sigma=rng.uniform(5, 40, ns))
    hp["r"][0], hp["dr"][0], hp["mu"][0], hp["sigma"][0], hp["c"][0], hp["n"][0] = 15, 5, 30, 5, 0.1, 10
    hp["r"][1], hp["dr"][1] = 20.0, 19.0      # wide prior: radii from 1 to 39 degrees
    res = dict(t=t, flux=f, flux_norm=fn, **hp)
    res["mean_ylm"] = np.zeros((ns, 256))
    res["cov_ylm_diag"] = np.zeros((ns, 256))
    res["cov_ylm_row6"] = np.zeros((ns, 256))
    res["cov_ylm_fro"] = np.zeros(ns)
    for m in (0, 1):
        for n_ in (0, 1):
            res["lnlike_m%d_n%d" % (m, n_)] = np.zeros(ns)
    for s in range(ns):
        kw = {k: float(hp[k][s]) for k in hp}
        for m in (0, 1):
            for n_ in (0, 1):
                g = SP(ydeg=15, marginalize_over_inclination=bool(m), normalized=bool(n_), **kw)
                res["lnlike_m%d_n%d" % (m, n_)][s] = float(
                    g.log_likelihood(t, fn if n_ else f, 1e-6, i=60.0, p=1.0, u=U_LD))
        res["mean_ylm"][s] = ev(g.mean_ylm)
        cy = ev(g.cov_ylm)
        res["cov_ylm_diag"][s] = np.diag(cy)
        res["cov_ylm_row6"][s] = cy[6]
        res["cov_ylm_fro"][s] = np.linalg.norm(cy)
        print(s, kw, [res["lnlike_m%d_n%d" % (a, b)][s] for a in (0, 1) for b in (0, 1)], flush=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "size_dr.npz"), **res)


if __name__ == "__main__":
    main()
