"""
TEST INFRASTRUCTURE ONLY.  Round-2 golden fixtures, produced by the UNMODIFIED reference through
oracle/theano_stub (default BLAS threading of the build container, like every other fixture):

    make -C oracle ref && python -m oracle.gen_golden_r2 [ensemble] [long] [fullprior]

* ``ensemble_nt1000.npz``  -- configs[1]: joint log-likelihood of the 1024 synthetic light curves of
  ``bench.ensemble_flux`` (regenerated from the seed, not stored) sharing the fiducial
  hyperparameters, in all four marginalise x normalise modes, plus a per-point ``data_cov`` vector
  with baseline terms.  Pins the ``M >= 64`` branch of ``StarryProcess.log_likelihood``
  (one factorisation + ``spb_cholesky_solve_rows``), which round 1 only tested at kernel level.
* ``longbaseline_nt4096_r2.npz`` -- configs[3]: 16 hyperparameter draws at nt = 4096, limb darkened,
  conditional on i = 60 deg; unnormalised with one element made non positive-definite through its
  per-element baseline variance (-> -inf, sp.py:1186-1188), and normalised with one draw outside the
  validity range z > 0.023 (-> -inf, sp.py:1178-1183).
* ``bench_fullprior_seed4321.npz`` -- the first 64 draws of bench.py's second line item (the
  reference's full stability prior: r in [10, 45], c in [0, 1], n in [1, 50]), marginalised +
  normalised: mostly -inf by the z-rule, the pattern and the finite values are pinned.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import theano_stub  # noqa: E402

FID = dict(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0)
U_LD = [0.4, 0.26]


def ev(x):
    return np.array(x.eval() if hasattr(x, "eval") else x, dtype=np.float64)


def ensemble(SP):
    M = 1024
    res = dict(M=M, seed=77, hyper=np.array([FID[k] for k in ("r", "mu", "sigma", "c", "n")]))
    dvec = 1e-6 * (1 + 0.5 * np.sin(np.arange(bench.NT)))
    res["data_cov_vec"] = dvec
    for norm in (False, True):
        t, f = bench.ensemble_flux(M, normalized=norm)
        res["flux_checksum_n%d" % norm] = float(np.sum(f * np.cos(np.arange(f.size)).reshape(f.shape)))
        for marg in (False, True):
            g = SP(ydeg=15, marginalize_over_inclination=marg, normalized=norm, **FID)
            key = "m%d_n%d" % (marg, norm)
            res["lnlike_" + key] = float(g.log_likelihood(t, f, 1e-6, i=60.0, p=1.0, u=U_LD))
            res["lnlike_dvec_" + key] = float(g.log_likelihood(
                t, f, dvec, i=60.0, p=1.0, u=U_LD, baseline_mean=1e-4, baseline_var=1e-5))
            # the first 64 curves alone (M == 64 is the smallest size that takes the branch)
            res["lnlike_64_" + key] = float(g.log_likelihood(t, f[:64], 1e-6, i=60.0, p=1.0, u=U_LD))
            print("ensemble", key, res["lnlike_" + key], res["lnlike_dvec_" + key], flush=True)
    np.savez_compressed(os.path.join(OUT, "ensemble_nt1000.npz"), **res)


def long_baseline(SP):
    nt = 4096
    old = np.load(os.path.join(OUT, "longbaseline_nt4096.npz"))
    t, f = old["t"], old["flux"]            # the round-1 light curve (GP draw + noise), reused
    fn = (1 + f) / np.mean(1 + f) - 1
    ns = 16
    rng = np.random.default_rng(40)
    hp = dict(r=rng.uniform(10, 30, ns), c=rng.uniform(0.01, 0.15, ns), n=rng.uniform(1, 12, ns),
              mu=rng.uniform(0, 85, ns), sigma=rng.uniform(5, 40, ns))
    hp["c"][11], hp["n"][11] = 0.6, 30.0    # normalised: z > 0.023 -> -inf
    bvar = np.zeros(ns)
    bvar[5] = -1e-3                          # K + bvar 1 1^T indefinite -> Cholesky fails -> -inf
    res = dict(t=t, flux=f, flux_norm=fn, u=np.array(U_LD), baseline_var=bvar, **hp)
    res["lnlike_n0"] = np.zeros(ns)
    res["lnlike_n1"] = np.zeros(ns)
    res["z_n1"] = np.zeros(ns)
    for s in range(ns):
        kw = {k: hp[k][s] for k in hp}
        g = SP(ydeg=15, marginalize_over_inclination=False, normalized=False, **kw)
        res["lnlike_n0"][s] = float(g.log_likelihood(t, f, 1e-6, i=60.0, p=1.0, u=U_LD,
                                                     baseline_var=bvar[s]))
        g = SP(ydeg=15, marginalize_over_inclination=False, normalized=True, **kw)
        res["lnlike_n1"][s] = float(g.log_likelihood(t, fn, 1e-6, i=60.0, p=1.0, u=U_LD))
        res["z_n1"][s] = float(ev(g._z))
        print("nt4096", s, res["lnlike_n0"][s], res["lnlike_n1"][s], res["z_n1"][s], flush=True)
    assert np.isneginf(res["lnlike_n0"][5]) and np.isneginf(res["lnlike_n1"][11])
    del res["t"], res["flux"]                # stored once, in longbaseline_nt4096.npz
    del res["flux_norm"]
    np.savez_compressed(os.path.join(OUT, "longbaseline_nt4096_r2.npz"), **res)


_SP = None


def _one_full(s):
    global _SP
    if _SP is None:
        _SP = theano_stub.import_reference()
    hp, t, flux, _ = bench.synthetic_inputs(4096, seed=4321, prior="full")
    g = _SP.StarryProcess(ydeg=15, marginalize_over_inclination=True, normalized=True,
                          r=hp["r"][s], mu=hp["mu"][s], sigma=hp["sigma"][s], c=hp["c"][s],
                          n=hp["n"][s])
    ll = float(g.log_likelihood(t, flux, 1e-6, i=60.0, p=1.0, u=bench.U_LD))
    return ll, float(ev(g._z))


def full_prior():
    import multiprocessing as mp

    NS = 64
    hp, t, flux, _ = bench.synthetic_inputs(4096, seed=4321, prior="full")
    res = {k: v[:NS].copy() for k, v in hp.items()}
    with mp.get_context("spawn").Pool(4) as pool:
        out = pool.map(_one_full, list(range(NS)), chunksize=4)
    res["lnlike_m1_n1"] = np.array([o[0] for o in out])
    res["z"] = np.array([o[1] for o in out])
    res["seed"] = 4321
    print("full prior: %d of %d finite" % (np.isfinite(res["lnlike_m1_n1"]).sum(), NS))
    np.savez_compressed(os.path.join(OUT, "bench_fullprior_seed4321.npz"), **res)


def main():
    which = sys.argv[1:] or ["ensemble", "long", "fullprior"]
    SP = theano_stub.import_reference().StarryProcess
    if "ensemble" in which:
        ensemble(SP)
    if "long" in which:
        long_baseline(SP)
    if "fullprior" in which:
        full_prior()


if __name__ == "__main__":
    main()
