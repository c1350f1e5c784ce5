"""
TEST INFRASTRUCTURE ONLY.  Golden fixtures for SURVEY.md section 8(f) rank 3 -- time-variable
surfaces: the temporal kernel Hadamard-multiplied into the flux covariance (temporal.py:8-16;
sp.py:697-698, 893-894) and ``sample_ylm(t)`` (sp.py:510-516, ops/sample.py:24-33) -- produced by
the UNMODIFIED reference package through ``oracle/theano_stub``.

    make -C oracle ref && python -m oracle.gen_golden_temporal
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
SEED = 5


def ev(x):
    return np.array(x.eval() if hasattr(x, "eval") else x, dtype=np.float64)


def main():
    sp = theano_stub.import_reference()
    SP = sp.StarryProcess
    g = np.load(os.path.join(OUT, "fiducial_nt1000.npz"))
    nt, nts = 300, 90
    t = np.ascontiguousarray(g["t"][:nt])
    ts = np.linspace(0.1, 1.4, nts)
    out = dict(t=t, t_sample=ts, flux=g["flux"][:nt], flux_norm=g["flux_norm"][:nt],
               u_ld=np.array(U_LD), hyper=np.array([FID[k] for k in ("r", "mu", "sigma", "c", "n")]))
    kernels = dict(matern32=(sp.Matern32Kernel, 0.7), expsq=(sp.ExpSquaredKernel, 0.05))
    for kname, (kern, tau) in kernels.items():
        out["tau_" + kname] = tau
        for marg in (False, True):
            for norm in (False, True):
                key = "%s_m%d_n%d" % (kname, marg, norm)
                gp = SP(ydeg=15, marginalize_over_inclination=marg, normalized=norm, tau=tau,
                        temporal_kernel=kern, seed=SEED, **FID)
                K = ev(gp.cov(t, i=60.0, p=1.0, u=U_LD))
                out["Krow0_" + key] = K[0].copy()
                out["Krow150_" + key] = K[150].copy()
                out["Kdiag_" + key] = np.diag(K).copy()
                f = out["flux_norm"] if norm else out["flux"]
                out["lnlike_" + key] = float(ev(gp.log_likelihood(t, f, 1e-6, i=60.0, p=1.0,
                                                                  u=U_LD)))
                if not norm:
                    mu, Kp = gp.predict(t, f, 1e-6, t_sample=ts, i=60.0, p=1.0, u=U_LD,
                                        baseline_var=1e-5)
                    out["pred_mu_" + key], out["pred_Kdiag_" + key] = ev(mu), np.diag(ev(Kp)).copy()
                    out["pred_Krow0_" + key] = ev(Kp)[0].copy()
    # ---- sample_ylm(t): y_i = L_t U_i L_y^T, U ~ (nsamples, nt, nylm)
    nty, ns = 12, 2
    ty = np.linspace(0, 1, nty)
    gp = SP(ydeg=15, tau=0.7, seed=SEED, **FID)
    gp.random._rng = np.random.RandomState(SEED)
    U = np.random.RandomState(SEED).normal(size=(ns, nty, 256))
    y = ev(gp.sample_ylm(t=ty, nsamples=ns))
    out["ylm_t"] = ty
    out["ylm_U"] = U
    out["ylm_y"] = y
    out["cov_ylm"] = ev(gp.cov_ylm)
    np.savez_compressed(os.path.join(OUT, "temporal_nt300.npz"), **out)
    for k, v in out.items():
        print(k, np.shape(v), float(np.abs(v).max()))


if __name__ == "__main__":
    main()
