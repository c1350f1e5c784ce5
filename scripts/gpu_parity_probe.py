"""GPU-box development probe: per-sample lnlike of the bench sweep against the live oracle, with the
stage-by-stage differences of the worst sample.  Writes gpurun_out/parity_probe.npz."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import starry_process_b200 as spb  # noqa: E402
from oracle import sp_oracle as so  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
hp, t, flux, fens = bench.synthetic_inputs(4096, seed=1234)
native = "ref" if so.ref_available(15, 2) else "port"
U = [0.4, 0.26]
out = {}
for marg in (True, False):
    for norm in (True, False):
        gp = spb.StarryProcess(marginalize_over_inclination=marg, normalized=norm,
                               **{k: torch.tensor(v[:N]) for k, v in hp.items()})
        ll = gp.log_likelihood(t, flux, 1e-6, i=60.0, p=1.0, u=U).cpu().numpy()
        ref = np.zeros(N)
        zs = np.zeros(N)
        for s in range(N):
            o = so.OracleProcess(r=hp["r"][s], mu=hp["mu"][s], sigma=hp["sigma"][s], c=hp["c"][s],
                                 n=hp["n"][s], native=native, marginalize_over_inclination=marg,
                                 normalized=norm)
            ref[s] = o.log_likelihood(t, flux, 1e-6, i=60.0, p=1.0, u=U)
            zs[s] = o.z if o.z is not None else 0.0
        fin = np.isfinite(ref) & np.isfinite(ll)
        err = np.zeros(N)
        err[fin] = np.abs(ll[fin] - ref[fin]) / np.abs(ref[fin])
        w = int(np.argmax(err))
        print("marg=%d norm=%d: max rel err %.3e at sample %d (r=%.2f mu=%.2f sigma=%.2f c=%.3f n=%.2f) "
              "median %.2e; -inf pattern equal: %s" % (
                  marg, norm, err[w], w, hp["r"][w], hp["mu"][w], hp["sigma"][w], hp["c"][w],
                  hp["n"][w], np.median(err), np.array_equal(np.isneginf(ll), np.isneginf(ref))),
              flush=True)
        key = "m%d_n%d" % (marg, norm)
        out["ll_" + key] = ll
        out["ref_" + key] = ref
        out["z_gpu_" + key] = gp._z.cpu().numpy()
        out["z_ref_" + key] = zs
        # stage diffs for the worst sample
        o = so.OracleProcess(r=hp["r"][w], mu=hp["mu"][w], sigma=hp["sigma"][w], c=hp["c"][w],
                             n=hp["n"][w], native=native, marginalize_over_inclination=marg,
                             normalized=norm)
        mu_g = gp.mean_ylm[w].cpu().numpy()
        cov_g = gp.cov_ylm[w].cpu().numpy()
        print("   mean_ylm maxdiff %.2e (scale %.2e)  cov_ylm maxdiff %.2e (scale %.2e)" % (
            np.abs(mu_g - o.mean_ylm).max(), np.abs(o.mean_ylm).max(),
            np.abs(cov_g - o.cov_ylm).max(), np.abs(o.cov_ylm).max()))
        gp1 = spb.StarryProcess(marginalize_over_inclination=marg, normalized=norm,
                                **{k: float(v[w]) for k, v in hp.items()})
        Kg = gp1.cov(t, i=60.0, p=1.0, u=U).cpu().numpy()
        Ko = o.cov(t, i=60.0, p=1.0, u=U)
        print("   cov(t) maxdiff %.2e (scale %.2e);  z gpu %.12e ref %.12e" % (
            np.abs(Kg - Ko).max(), np.abs(Ko).max(), float(gp1._z[0]), o.z if o.z else 0.0))
        # which stage: feed the oracle's Ylm moments through the GPU flux path and vice versa
        o2 = so.OracleProcess(r=hp["r"][w], mu=hp["mu"][w], sigma=hp["sigma"][w], c=hp["c"][w],
                              n=hp["n"][w], native=native, marginalize_over_inclination=marg,
                              normalized=norm)
        o2.mean_ylm = mu_g.copy()
        o2.cov_ylm = cov_g.copy()
        o2.ez = o2._dotRx(o2.mean_ylm.reshape(1, -1), o2._rx90).T
        mom2y = np.ascontiguousarray(o2.cov_ylm + np.outer(o2.mean_ylm, o2.mean_ylm))
        tmp = np.ascontiguousarray(o2._dotRx(mom2y, o2._rx90).T)
        o2.Ez = o2._dotRx(tmp, o2._rx90)
        ll2 = o2.log_likelihood(t, flux, 1e-6, i=60.0, p=1.0, u=U)
        print("   oracle flux path on GPU Ylm moments: lnlike %.10f  (gpu %.10f, oracle %.10f) -> "
              "moments explain %.2e, flux/cholesky path %.2e" % (
                  ll2, ll[w], ref[w], abs(ll2 - ref[w]) / abs(ref[w]), abs(ll2 - ll[w]) / abs(ref[w])))
        out["worst_" + key] = w
        out["cov_g_" + key] = cov_g
        out["cov_o_" + key] = o.cov_ylm
np.savez(os.path.join(ROOT, "gpurun_out", "parity_probe.npz"), **out)
