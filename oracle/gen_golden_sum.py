"""
TEST INFRASTRUCTURE ONLY.  Golden fixtures for ``StarryProcessSum`` (sp.py:1190-1198, 1335-1400:
``sp1 + sp2`` adds the Ylm means and covariances) and ``StarryProcess.flux`` (sp.py:1237-1283),
produced by the UNMODIFIED reference package through ``oracle/theano_stub``.

    make -C oracle ref && python -m oracle.gen_golden_sum
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import theano_stub  # noqa: E402

FID = dict(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0)
SEC = dict(r=20.0, mu=60.0, sigma=10.0, c=0.05, n=5.0)
U_LD = [0.4, 0.26]
SEED = 21


def ev(x):
    return np.array(x.eval() if hasattr(x, "eval") else x, dtype=np.float64)


def main():
    sp = theano_stub.import_reference()
    SP = sp.StarryProcess
    g = np.load(os.path.join(OUT, "fiducial_nt1000.npz"))
    nt = 300
    t = np.ascontiguousarray(g["t"][:nt])
    out = dict(t=t, flux=g["flux"][:nt], flux_norm=g["flux_norm"][:nt],
               first=np.array([FID[k] for k in ("r", "mu", "sigma", "c", "n")]),
               second=np.array([SEC[k] for k in ("r", "mu", "sigma", "c", "n")]))
    for marg in (False, True):
        for norm in (False, True):
            key = "m%d_n%d" % (marg, norm)
            kw = dict(ydeg=15, marginalize_over_inclination=marg, normalized=norm, seed=SEED)
            gp = SP(**kw, **FID) + SP(**kw, **SEC)
            f = out["flux_norm"] if norm else out["flux"]
            out["lnlike_" + key] = float(ev(gp.log_likelihood(t, f, 1e-6, i=60.0, p=1.0, u=U_LD)))
            K = ev(gp.cov(t, i=60.0, p=1.0, u=U_LD))
            out["Krow100_" + key] = K[100].copy()
    gp = SP(ydeg=15, normalized=True, seed=SEED, **FID) + SP(ydeg=15, normalized=True, seed=SEED, **SEC)
    out["mean_ylm"] = ev(gp.mean_ylm)
    out["cov_ylm"] = ev(gp.cov_ylm)
    gp.random._rng = np.random.RandomState(SEED)
    U = np.random.RandomState(SEED).normal(size=(256, 3))
    y = ev(gp.sample_ylm(nsamples=3))
    out["ylm_U"], out["ylm_y"] = U, y
    # flux of given Ylm vectors (normalised and not)
    out["flux_of_y_n1"] = ev(gp.flux(y, t, i=60.0, p=1.0, u=U_LD))
    gpn = SP(ydeg=15, normalized=False, seed=SEED, **FID)
    out["flux_of_y_n0"] = ev(gpn.flux(y, t, i=60.0, p=1.0, u=U_LD))
    np.savez_compressed(os.path.join(OUT, "sum_flux_nt300.npz"), **out)
    for k, v in out.items():
        print(k, np.shape(v), float(np.abs(v).max()))


if __name__ == "__main__":
    main()
