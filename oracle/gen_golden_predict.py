"""
TEST INFRASTRUCTURE ONLY.  Golden fixtures for SURVEY.md section 8(f) rank 2 -- ``sample``,
``predict``, ``sample_conditional`` and ``sample_ylm_conditional`` (sp.py:518-641, 729-765,
767-1002) -- produced by the UNMODIFIED reference package through ``oracle/theano_stub``.

    make -C oracle ref && python -m oracle.gen_golden_predict

The standard-normal draws the reference takes from its ``RandomStream`` are recorded by replaying
the same ``RandomState`` (the stub's stream is ``numpy.random.RandomState(seed).normal``), so the
CUDA path can be handed identical draws.

The reference's ``sample_conditional`` (sp.py:985-1002) refers to an undefined name ``ts`` and
raises ``NameError`` as shipped; its golden values are therefore assembled here from the
reference's own ``predict`` output exactly as that method's body prescribes
(``cho_factor(K + eps I)``, ``mu + L U``), with ``ts = t_sample``.
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
U_LD = [0.4, 0.26]
SEED = 11


def ev(x):
    return np.array(x.eval() if hasattr(x, "eval") else x, dtype=np.float64)


def main():
    sp = theano_stub.import_reference()
    SP = sp.StarryProcess
    g = np.load(os.path.join(OUT, "fiducial_nt1000.npz"))
    nt, nts, ns = 200, 120, 3
    t = np.ascontiguousarray(g["t"][:nt])
    flux = np.ascontiguousarray(g["flux"][:nt])
    ts = np.linspace(0.3, 1.9, nts)
    dvec = 1e-6 * (1 + 0.5 * np.sin(np.arange(nt)))
    out = dict(t=t, flux=flux, t_sample=ts, data_cov_vec=dvec, u_ld=np.array(U_LD),
               hyper=np.array([FID[k] for k in ("r", "mu", "sigma", "c", "n")]), eps=1e-8)

    for marg in (False, True):
        tag = "m%d" % marg
        gp = SP(ydeg=15, marginalize_over_inclination=marg, normalized=False, seed=SEED, **FID)
        # ---- sample (sp.py:729-765): U ~ (nt, nsamples) from the stream
        gp.random._rng = np.random.RandomState(SEED)
        U = np.random.RandomState(SEED).normal(size=(nt, ns))
        s = ev(gp.sample(t, i=60.0, p=1.0, u=U_LD, nsamples=ns, eps=1e-8))
        out["sample_U_" + tag] = U
        out["sample_" + tag] = s
        # ---- predict (sp.py:767-922), on t and on t_sample, scalar / vector data_cov,
        #      with and without baseline variance
        mu, K = gp.predict(t, flux, 1e-6, i=60.0, p=1.0, u=U_LD)
        out["pred_mu_" + tag], out["pred_K_" + tag] = ev(mu), ev(K)
        mu, K = gp.predict(t, flux, dvec, t_sample=ts, i=60.0, p=1.0, u=U_LD,
                           baseline_mean=1e-4, baseline_var=1e-5)
        out["pred_ts_mu_" + tag], out["pred_ts_K_" + tag] = ev(mu), ev(K)
        # ---- sample_conditional (see the module docstring)
        Uc = np.random.RandomState(SEED + 1).normal(size=(nts, ns))
        Kc = out["pred_ts_K_" + tag]
        Lc = np.linalg.cholesky(Kc + 1e-8 * np.eye(nts))
        out["cond_U_" + tag] = Uc
        out["cond_sample_" + tag] = (out["pred_ts_mu_" + tag][:, None] + Lc @ Uc).T

    # ---- sample_ylm_conditional (sp.py:518-641); the inclination is always used
    gp = SP(ydeg=15, marginalize_over_inclination=False, normalized=False, seed=SEED, **FID)
    gp.random._rng = np.random.RandomState(SEED + 2)
    Uy = np.random.RandomState(SEED + 2).normal(size=(256, ns))
    y = ev(gp.sample_ylm_conditional(t, flux, 1e-6, i=60.0, p=1.0, u=U_LD, baseline_mean=1e-4,
                                     baseline_var=1e-5, nsamples=ns))
    out["ylmc_U"] = Uy
    out["ylmc_y"] = y
    # the flux the conditional draws imply (a well-conditioned functional of y)
    A = ev(gp._flux.design_matrix(t, 60.0, 1.0, U_LD))
    out["ylmc_flux"] = y @ A.T
    np.savez_compressed(os.path.join(OUT, "predict_nt200.npz"), **out)
    for k, v in out.items():
        print(k, np.shape(v))


if __name__ == "__main__":
    main()
