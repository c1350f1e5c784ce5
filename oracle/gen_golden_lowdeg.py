"""
TEST INFRASTRUCTURE ONLY.  Golden fixture for spherical-harmonic degrees below 15 and for the
keyword options the reference threads through its integrals (sp.py:241-262), produced by the
UNMODIFIED reference through oracle/theano_stub:

    make -C oracle ref && python -m oracle.gen_golden_lowdeg

``lowdeg_options.npz``: ydeg = 5 and 10 (default options) and ydeg = 15 with epsy = 1e-10,
epsy15 = 1e-8, abmin = 1e-3 (one draw clamped by it), log_alpha_max = 8, log_beta_max = 9; for each,
4 hyperparameter draws: mean_ylm, cov_ylm (full), lnlike in the four marginalise x normalise modes on
an nt = 200 light curve, and a design-matrix row block.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import theano_stub  # noqa: E402

U_LD = [0.4, 0.26]
CASES = {
    "y5": dict(ydeg=5),
    "y10": dict(ydeg=10),
    "y15opt": dict(ydeg=15, epsy=1e-10, epsy15=1e-8, abmin=1e-3, log_alpha_max=8.0, log_beta_max=9.0),
}


def ev(x):
    return np.array(x.eval() if hasattr(x, "eval") else x, dtype=np.float64)


def main():
    SP = theano_stub.import_reference().StarryProcess
    rng = np.random.default_rng(31)
    ns = 4
    hp = dict(r=rng.uniform(12, 30, ns), a=rng.uniform(0.2, 0.7, ns), b=rng.uniform(0.1, 0.6, ns),
              c=rng.uniform(0.02, 0.15, ns), n=rng.uniform(1, 12, ns))
    hp["a"][3] = 2e-4          # below abmin = 1e-3 of the option case: clamped there
    t = np.linspace(0, 3, 200)
    f = 2e-3 * np.sin(2 * np.pi * t + 0.3) + 1e-3 * rng.standard_normal(200)
    fn = (1 + f) / np.mean(1 + f) - 1
    res = dict(t=t, flux=f, flux_norm=fn, **hp)
    for name, kw in CASES.items():
        ny = (kw["ydeg"] + 1) ** 2
        res[name + "_mean"] = np.zeros((ns, ny))
        res[name + "_cov"] = np.zeros((ns, ny, ny))
        for m in (0, 1):
            for n_ in (0, 1):
                res["%s_lnlike_m%d_n%d" % (name, m, n_)] = np.zeros(ns)
        for s in range(ns):
            h = {k: float(hp[k][s]) for k in hp}
            for m in (0, 1):
                for n_ in (0, 1):
                    g = SP(marginalize_over_inclination=bool(m), normalized=bool(n_), **h, **kw)
                    res["%s_lnlike_m%d_n%d" % (name, m, n_)][s] = float(
                        g.log_likelihood(t, fn if n_ else f, 1e-6, i=55.0, p=0.8, u=U_LD))
            res[name + "_mean"][s] = ev(g.mean_ylm)
            res[name + "_cov"][s] = ev(g.cov_ylm)
            print(name, s, [res["%s_lnlike_m%d_n%d" % (name, a, b)][s] for a in (0, 1) for b in (0, 1)],
                  flush=True)
        res[name + "_A"] = ev(g._flux.design_matrix(t[:7], 55.0, 0.8, U_LD))
    # keep the fixture small: the full degree-15 covariance is 2 MB; its diagonal suffices
    res["y15opt_cov_diag"] = np.diagonal(res.pop("y15opt_cov"), axis1=1, axis2=2).copy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lowdeg_options.npz"), **res)


if __name__ == "__main__":
    main()
