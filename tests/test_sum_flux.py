"""``StarryProcessSum`` (reference sp.py:1190-1198, 1335-1400) and ``StarryProcess.flux``
(sp.py:1237-1283): oracle (CPU) and CUDA path (-m gpu) against fixtures produced by the unmodified
reference (oracle/gen_golden_sum.py).  lnlike 1e-8 relative, covariance rows 1e-10 of the matrix
scale, flux of given Ylm vectors 1e-12."""
import numpy as np
import pytest

from conftest import FID, U_LD

SEC = dict(r=20.0, mu=60.0, sigma=10.0, c=0.05, n=5.0)
KW = dict(i=60.0, p=1.0, u=U_LD)


@pytest.mark.parametrize("marg", [False, True])
@pytest.mark.parametrize("norm", [False, True])
def test_oracle_sum_vs_reference(oracle, golden, marg, norm):
    g = golden("sum_flux_nt300.npz")
    key = "m%d_n%d" % (marg, norm)
    kw = dict(marginalize_over_inclination=marg, normalized=norm)
    o = oracle.OracleProcess(**kw, **FID) + oracle.OracleProcess(**kw, **SEC)
    f = g["flux_norm"] if norm else g["flux"]
    ll = o.log_likelihood(g["t"], f, 1e-6, **KW)
    assert abs(ll - g["lnlike_" + key]) <= 1e-12 * abs(g["lnlike_" + key])
    K = o.cov(g["t"], **KW)
    assert np.abs(K[100] - g["Krow100_" + key]).max() <= 1e-13 * np.abs(K).max()


def test_oracle_flux_and_sum_moments(oracle, golden):
    g = golden("sum_flux_nt300.npz")
    o = oracle.OracleProcess(**FID) + oracle.OracleProcess(**SEC)
    assert np.abs(o.mean_ylm - g["mean_ylm"]).max() <= 1e-16
    assert np.abs(o.cov_ylm - g["cov_ylm"]).max() <= 1e-15 * np.abs(g["cov_ylm"]).max()
    y = o.sample_ylm(g["ylm_U"])
    assert np.abs(y - g["ylm_y"]).max() <= 1e-9 * np.abs(g["ylm_y"]).max()
    F = o.flux(g["ylm_y"], g["t"], **KW)
    assert np.abs(F - g["flux_of_y_n1"]).max() <= 1e-13
    o0 = oracle.OracleProcess(normalized=False, **FID)
    assert np.abs(o0.flux(g["ylm_y"], g["t"], **KW) - g["flux_of_y_n0"]).max() <= 1e-13


@pytest.fixture(scope="module")
def spb():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import starry_process_b200 as m

    return m


@pytest.mark.gpu
@pytest.mark.parametrize("marg", [False, True])
@pytest.mark.parametrize("norm", [False, True])
def test_gpu_sum_vs_reference_golden(spb, golden, marg, norm):
    g = golden("sum_flux_nt300.npz")
    key = "m%d_n%d" % (marg, norm)
    kw = dict(marginalize_over_inclination=marg, normalized=norm)
    gp = spb.StarryProcess(**kw, **FID) + spb.StarryProcess(**kw, **SEC)
    assert isinstance(gp, spb.StarryProcessSum) and len(gp._children) == 2
    f = g["flux_norm"] if norm else g["flux"]
    ll = gp.log_likelihood(g["t"], f, 1e-6, **KW).item()
    assert abs(ll - g["lnlike_" + key]) <= 1e-8 * abs(g["lnlike_" + key])
    K = gp.cov(g["t"], **KW).cpu().numpy()
    assert np.abs(K[100] - g["Krow100_" + key]).max() <= 1e-10 * np.abs(K).max()


@pytest.mark.gpu
def test_gpu_sum_of_three_flux_and_errors(spb, golden):
    import torch

    g = golden("sum_flux_nt300.npz")
    a, b = spb.StarryProcess(**FID), spb.StarryProcess(**SEC)
    s3 = sum([a, b, spb.StarryProcess(r=15.0, mu=10.0, sigma=5.0, c=0.02, n=2.0)])   # __radd__
    assert len(s3._children) == 3
    assert bool(torch.isfinite(s3.log_likelihood(g["t"], g["flux_norm"], 1e-6, **KW)))
    assert float((s3.mean_ylm - (a + b).mean_ylm).abs().max()) > 0
    # flux of given Ylm vectors, normalised and not (sp.py:1237-1283)
    F1 = a.flux(g["ylm_y"], g["t"], **KW).cpu().numpy()
    assert F1.shape == (3, 300) and np.abs(F1 - g["flux_of_y_n1"]).max() <= 1e-12
    a0 = spb.StarryProcess(normalized=False, **FID)
    F0 = a0.flux(g["ylm_y"], g["t"], **KW).cpu().numpy()
    assert np.abs(F0 - g["flux_of_y_n0"]).max() <= 1e-12
    assert tuple(a0.flux(g["ylm_y"][0], g["t"], **KW).shape) == (300,)
    # time-variable surfaces: one Ylm vector per time
    gt = spb.StarryProcess(tau=0.5, normalized=False, **FID)
    yt = gt.sample_ylm(t=g["t"][:20], nsamples=2)
    Ft = gt.flux(yt, g["t"][:20], **KW)
    A = gt.design_matrix(g["t"][:20], **KW)
    assert float((Ft - torch.einsum("kn,skn->sk", A, yt)).abs().max()) <= 1e-15
    with pytest.raises(AssertionError):
        spb.StarryProcess(normalized=True, **FID) + spb.StarryProcess(normalized=False, **SEC)
    with pytest.raises(AssertionError):
        spb.StarryProcess(tau=1.0, **FID) + spb.StarryProcess(**SEC)
    with pytest.raises(NotImplementedError):
        (a + b).log_jac()
