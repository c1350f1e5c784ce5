"""
TEST INFRASTRUCTURE ONLY.  Generates the golden fixtures under ``tests/golden/`` by running the
UNMODIFIED reference package (``/root/reference/starry_process``) through ``oracle/theano_stub``
(eager NumPy stand-in for Theano; the reference's C++ ops come from ``oracle/_ref``).

    make -C oracle ref && python -m oracle.gen_golden

Run in the build container only (the GPU box has no /root/reference); the resulting ``.npz``
files are committed.  Inputs follow SURVEY.md section 8(d): ``t = linspace(0, 4, nt)``, ``p = 1``,
``data_cov = 1e-6``, fiducial ``r=10, mu=30, sigma=5, c=0.1, n=10``; flux = reference GP draw at the
fiducial point (conditional, i = 60 deg) + N(0, 1e-3^2), ``numpy.random.default_rng(0)``;
hyperparameter sweeps from the reference's own stability prior (joss/figures/stability.py:28-35).
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


def ev(x):
    return np.array(x.eval() if hasattr(x, "eval") else x, dtype=np.float64)


def draw_sweep(nsamp, seed=0, lowc=False):
    """stability.py prior; ``lowc`` shrinks contrast/number so that the normalised process stays
    inside its validity range z <= 0.023 (sp.py:1178-1183) instead of returning -inf."""
    rng = np.random.default_rng(seed)
    if lowc:
        return dict(
            r=rng.uniform(10, 30, nsamp), c=rng.uniform(0.01, 0.15, nsamp),
            n=rng.uniform(1, 12, nsamp), mu=rng.uniform(0, 85, nsamp),
            sigma=rng.uniform(5, 40, nsamp),
        )
    return dict(
        r=rng.uniform(10, 45, nsamp), c=rng.uniform(0, 1, nsamp), n=rng.uniform(1, 50, nsamp),
        mu=rng.uniform(0, 85, nsamp), sigma=rng.uniform(5, 40, nsamp),
    )


def main():
    os.makedirs(OUT, exist_ok=True)
    sp = theano_stub.import_reference()
    SP = sp.StarryProcess

    # ------------------------------------------------------------------ fiducial, nt = 1000
    nt = 1000
    t = np.linspace(0, 4, nt)
    gp = SP(ydeg=15, marginalize_over_inclination=False, normalized=False, **FID)
    mean_ylm = ev(gp.mean_ylm)
    cov_ylm = ev(gp.cov_ylm)
    cho_ylm = ev(gp.cho_cov_ylm)
    Kc = ev(gp.cov(t, i=60.0, p=1.0, u=[0.0, 0.0]))
    mc = ev(gp.mean(t, i=60.0, p=1.0, u=[0.0, 0.0]))
    rng = np.random.default_rng(0)
    Lk = np.linalg.cholesky(Kc + 1e-12 * np.eye(nt))
    nens = 8
    flux_ens = (mc[None, :] + (Lk @ rng.standard_normal((nt, nens))).T
                + 1e-3 * rng.standard_normal((nens, nt)))
    flux = flux_ens[0].copy()
    # normalized light curves are mean-normalised, zero baseline (sp.py:1067-1075)
    flux_norm = (1 + flux) / np.mean(1 + flux) - 1
    flux_ens_norm = (1 + flux_ens) / np.mean(1 + flux_ens, axis=1, keepdims=True) - 1

    out = dict(t=t, flux=flux, flux_norm=flux_norm, flux_ens=flux_ens,
               flux_ens_norm=flux_ens_norm, mean_ylm=mean_ylm, cov_ylm=cov_ylm,
               cho_ylm_diag=np.diag(cho_ylm).copy(), data_cov=1e-6,
               hyper=np.array([FID[k] for k in ("r", "mu", "sigma", "c", "n")]))
    a, b = sp.gauss2beta(FID["mu"], FID["sigma"])
    out["ab"] = np.array([a, b])
    for marg in (False, True):
        for norm in (False, True):
            g = SP(ydeg=15, marginalize_over_inclination=marg, normalized=norm, **FID)
            for uname, u in (("u0", [0.0, 0.0]), ("uld", U_LD)):
                key = "m%d_n%d_%s" % (marg, norm, uname)
                f = flux_norm if norm else flux
                fe = flux_ens_norm if norm else flux_ens
                out["lnlike_" + key] = float(g.log_likelihood(t, f, 1e-6, i=60.0, p=1.0, u=u))
                out["lnlike_ens_" + key] = float(g.log_likelihood(t, fe, 1e-6, i=60.0, p=1.0, u=u))
                K = ev(g.cov(t, i=60.0, p=1.0, u=u))
                out["Krow0_" + key] = K[0].copy()
                out["Krow500_" + key] = K[500].copy()
                out["Kdiag_" + key] = np.diag(K).copy()
                out["Ksum_" + key] = float(K.sum())
                out["gpmean_" + key] = float(ev(g.mean(t, i=60.0, p=1.0, u=u))[0])
                if norm:
                    out["z_" + key] = float(ev(g._z))
                print(key, out["lnlike_" + key], out["lnlike_ens_" + key])
            # extra data_cov / baseline forms (sp.py:1135-1151)
            key = "m%d_n%d" % (marg, norm)
            f = flux_norm if norm else flux
            dvec = 1e-6 * (1 + 0.5 * np.sin(np.arange(nt)))
            out["lnlike_dvec_" + key] = float(
                g.log_likelihood(t, f, dvec, i=60.0, p=1.0, u=U_LD, baseline_mean=1e-4,
                                 baseline_var=1e-5))
    out["data_cov_vec"] = 1e-6 * (1 + 0.5 * np.sin(np.arange(nt)))
    np.savez_compressed(os.path.join(OUT, "fiducial_nt1000.npz"), **out)

    # ------------------------------------------------------------------ sample_ylm (explicit u)
    Un = np.random.default_rng(1).standard_normal((256, 4))
    y = (mean_ylm[:, None] + cho_ylm @ Un).T
    np.savez_compressed(os.path.join(OUT, "sample_ylm.npz"), unit_normals=Un, y=y,
                        hyper=out["hyper"])

    # ------------------------------------------------------------------ hyperparameter sweep
    for sweep_name, lowc, seed in (("sweep_nt1000", False, 0), ("sweep_lowc_nt1000", True, 7)):
        nsamp = 24
        sw = draw_sweep(nsamp, seed=seed, lowc=lowc)
        res = {k: v for k, v in sw.items()}
        keys = [(m, n_) for m in (False, True) for n_ in (False, True)]
        for (m, n_) in keys:
            res["lnlike_m%d_n%d" % (m, n_)] = np.zeros(nsamp)
        res["mean_ylm"] = np.zeros((nsamp, 256))
        res["cov_ylm_diag"] = np.zeros((nsamp, 256))
        res["cov_ylm_fro"] = np.zeros(nsamp)
        res["cov_ylm_row6"] = np.zeros((nsamp, 256))
        for s in range(nsamp):
            hp = dict(r=sw["r"][s], mu=sw["mu"][s], sigma=sw["sigma"][s], c=sw["c"][s], n=sw["n"][s])
            for (m, n_) in keys:
                g = SP(ydeg=15, marginalize_over_inclination=m, normalized=n_, **hp)
                f = flux_norm if n_ else flux
                res["lnlike_m%d_n%d" % (m, n_)][s] = float(
                    g.log_likelihood(t, f, 1e-6, i=60.0, p=1.0, u=U_LD))
            res["mean_ylm"][s] = ev(g.mean_ylm)
            cy = ev(g.cov_ylm)
            res["cov_ylm_diag"][s] = np.diag(cy)
            res["cov_ylm_fro"][s] = np.linalg.norm(cy)
            res["cov_ylm_row6"][s] = cy[6]
            print(sweep_name, s, hp, [res["lnlike_m%d_n%d" % k][s] for k in keys])
        res["t"] = t
        np.savez_compressed(os.path.join(OUT, sweep_name + ".npz"), **res)

    # ------------------------------------------------------------------ long baseline, LD
    nt4 = 4096
    t4 = np.linspace(0, 16, nt4)
    rng = np.random.default_rng(4)
    g = SP(ydeg=15, marginalize_over_inclination=False, normalized=False, **FID)
    K4 = ev(g.cov(t4, i=60.0, p=1.0, u=U_LD))
    m4 = ev(g.mean(t4, i=60.0, p=1.0, u=U_LD))
    f4 = (m4 + np.linalg.cholesky(K4 + 1e-12 * np.eye(nt4)) @ rng.standard_normal(nt4)
          + 1e-3 * rng.standard_normal(nt4))
    res4 = dict(t=t4, flux=f4, u=np.array(U_LD))
    sw4 = draw_sweep(3, seed=4)
    res4.update({k: v for k, v in sw4.items()})
    res4["lnlike"] = np.zeros(3)
    for s in range(3):
        hp = dict(r=sw4["r"][s], mu=sw4["mu"][s], sigma=sw4["sigma"][s], c=sw4["c"][s],
                  n=sw4["n"][s])
        g = SP(ydeg=15, marginalize_over_inclination=False, normalized=False, **hp)
        res4["lnlike"][s] = float(g.log_likelihood(t4, f4, 1e-6, i=60.0, p=1.0, u=U_LD))
        print("nt4096", s, res4["lnlike"][s])
    np.savez_compressed(os.path.join(OUT, "longbaseline_nt4096.npz"), **res4)

    # ------------------------------------------------------------------ design matrix
    gpd = SP(ydeg=15, marginalize_over_inclination=False, normalized=False, **FID)
    td = np.array([0.0, 0.013, 0.1, 0.25, 0.37, 0.5, 0.77, 0.999, 1.0, 3.21])
    incs = np.array([0.0, 15.0, 60.0, 89.0, 90.0])
    Ad = np.zeros((2, len(incs), len(td), 256))
    for k, u in enumerate(([0.0, 0.0], U_LD)):
        for j, inc in enumerate(incs):
            Ad[k, j] = ev(gpd._flux.design_matrix(td, inc, 1.0, u))
    rta1 = np.stack([ev(gpd._flux._rTA1) for _ in range(1)])
    np.savez_compressed(os.path.join(OUT, "design_matrix_ref.npz"), t=td, incs=incs, A=Ad,
                        u=np.array(U_LD), p=1.0, rTA1L_last=rta1)

    # the reference repository's one golden numeric artefact (starry design matrices,
    # app/design.py:58-95), sub-sampled in phase to keep the fixture small
    AF = np.load(os.path.join(theano_stub.REFERENCE_ROOT, "starry_process", "app", "data",
                              "A_F15-300.npz"))["A_F"]
    theta_deg = np.linspace(0, 360, 300) * 2
    sel = np.arange(0, 300, 13)
    np.savez_compressed(os.path.join(OUT, "design_matrix_AF15.npz"), A_F=AF[:, sel, :],
                        theta_deg=theta_deg[sel], incs=np.array([15, 30, 45, 60, 75, 90.0]))

    # ------------------------------------------------------------------ latitude integrals
    lib = theano_stub.ref_lib(15, 2)
    ab = np.array([[53.59815003, 3.22602246], [1.0, 0.5], [22026.4657948, 22026.4657948],
                   [3.3, 0.9], [1.00001, 17.0], [400.0, 0.5000001]])
    qs = np.zeros((len(ab), 256))
    Qsub = np.zeros((len(ab), 256, 16))
    Qtr = np.zeros(len(ab))
    for k, (al, be) in enumerate(ab):
        q = np.empty(256)
        Q = np.empty((256, 256))
        lib.ref_latitude(al, be, theano_stub._ptr(q), theano_stub._ptr(Q))
        qs[k] = q
        Qsub[k] = Q[:, ::16]
        Qtr[k] = np.trace(Q)
    np.savez_compressed(os.path.join(OUT, "latitude_integrals.npz"), alpha_beta=ab, q=qs,
                        Q_cols_0_16_32=Qsub, Q_trace=Qtr)
    print("done; fixtures in", OUT)
    os.system("ls -la %s" % OUT)


if __name__ == "__main__":
    main()
