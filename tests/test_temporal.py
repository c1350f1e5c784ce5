"""SURVEY.md section 8(f) rank 3 -- time-variable surfaces (reference temporal.py:8-16,
sp.py:225-232, 697-698, 893-894, 510-516, ops/sample.py:24-33).

CPU: oracle restatement against fixtures produced by the unmodified reference
(oracle/gen_golden_temporal.py).  GPU (-m gpu): the CUDA path against the same fixtures:
lnlike 1e-8 relative (all four marginalise / normalise modes, both kernels), covariance rows 1e-10
of the matrix scale, predictive mean 1e-8, ``sample_ylm(t)`` 1e-8 given the reference's cov_ylm.
"""
import numpy as np
import pytest

from conftest import FID, U_LD

KW = dict(i=60.0, p=1.0, u=U_LD)
KERNELS = ("matern32", "expsq")


def _oracle_kernel(oracle, name):
    return oracle.Matern32Kernel if name == "matern32" else oracle.ExpSquaredKernel


@pytest.mark.parametrize("kname", KERNELS)
@pytest.mark.parametrize("marg", [False, True])
@pytest.mark.parametrize("norm", [False, True])
def test_oracle_temporal_vs_reference(oracle, golden, kname, marg, norm):
    g = golden("temporal_nt300.npz")
    key = "%s_m%d_n%d" % (kname, marg, norm)
    o = oracle.OracleProcess(marginalize_over_inclination=marg, normalized=norm,
                             tau=float(g["tau_" + kname]),
                             temporal_kernel=_oracle_kernel(oracle, kname), **FID)
    f = g["flux_norm"] if norm else g["flux"]
    ll = o.log_likelihood(g["t"], f, 1e-6, **KW)
    assert abs(ll - g["lnlike_" + key]) <= 1e-12 * abs(g["lnlike_" + key])
    K = o.cov(g["t"], **KW)
    assert np.abs(K[150] - g["Krow150_" + key]).max() <= 1e-13 * np.abs(K).max()
    if not norm:
        mu, Kp = o.predict(g["t"], f, 1e-6, t_sample=g["t_sample"], baseline_var=1e-5, **KW)
        assert np.abs(mu - g["pred_mu_" + key]).max() <= 1e-11 * np.abs(mu).max()


def test_oracle_sample_ylm_temporal(oracle, golden):
    g = golden("temporal_nt300.npz")
    o = oracle.OracleProcess(tau=0.7, **FID)
    y = o.sample_ylm_temporal(g["ylm_t"], g["ylm_U"])
    assert np.abs(y - g["ylm_y"]).max() <= 1e-12 * np.abs(g["ylm_y"]).max()
    with pytest.raises(NotImplementedError):
        o2 = oracle.OracleProcess(tau=0.7, normalized=False, **FID)
        o2.sample_ylm_conditional(g["t"], g["flux"], 1e-6, np.zeros((256, 1)))


# ------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def spb():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import starry_process_b200 as m

    return m


def _kernel(spb, name):
    return spb.Matern32Kernel if name == "matern32" else spb.ExpSquaredKernel


@pytest.mark.gpu
@pytest.mark.parametrize("kname", KERNELS)
@pytest.mark.parametrize("marg", [False, True])
@pytest.mark.parametrize("norm", [False, True])
def test_gpu_temporal_vs_reference_golden(spb, golden, kname, marg, norm):
    g = golden("temporal_nt300.npz")
    key = "%s_m%d_n%d" % (kname, marg, norm)
    gp = spb.StarryProcess(marginalize_over_inclination=marg, normalized=norm,
                           tau=float(g["tau_" + kname]), temporal_kernel=_kernel(spb, kname), **FID)
    f = g["flux_norm"] if norm else g["flux"]
    ll = gp.log_likelihood(g["t"], f, 1e-6, **KW).item()
    assert abs(ll - g["lnlike_" + key]) <= 1e-8 * abs(g["lnlike_" + key])
    K = gp.cov(g["t"], **KW).cpu().numpy()
    scale = np.abs(K).max()
    assert np.abs(K[0] - g["Krow0_" + key]).max() <= 1e-10 * scale
    assert np.abs(K[150] - g["Krow150_" + key]).max() <= 1e-10 * scale
    assert np.abs(np.diag(K) - g["Kdiag_" + key]).max() <= 1e-10 * scale
    if not norm:
        mu, Kp = gp.predict(g["t"], f, 1e-6, t_sample=g["t_sample"], baseline_var=1e-5, **KW)
        mu, Kp = mu.cpu().numpy(), Kp.cpu().numpy()
        assert np.abs(mu - g["pred_mu_" + key]).max() <= 1e-8 * np.abs(mu).max()
        assert np.abs(np.diag(Kp) - g["pred_Kdiag_" + key]).max() <= 1e-8 * scale
        assert np.abs(Kp[0] - g["pred_Krow0_" + key]).max() <= 1e-8 * scale


@pytest.mark.gpu
def test_gpu_temporal_batched_tau(spb, golden):
    """``tau`` may be a (B,) tensor like the other hyperparameters."""
    g = golden("temporal_nt300.npz")
    gp = spb.StarryProcess(tau=[0.7, 0.3], **FID)
    ll = gp.log_likelihood(g["t"], g["flux_norm"], 1e-6, **KW).cpu().numpy()
    assert ll.shape == (2,)
    assert abs(ll[0] - g["lnlike_matern32_m1_n1"]) <= 1e-8 * abs(g["lnlike_matern32_m1_n1"])
    assert ll[1] != ll[0]
    with pytest.raises(ValueError):
        spb.StarryProcess(tau=-1.0, **FID)
    with pytest.raises(NotImplementedError):
        spb.StarryProcess(tau=0.5, temporal_kernel=lambda a, b, c: None, **FID)


@pytest.mark.gpu
def test_gpu_sample_ylm_temporal(spb, golden):
    import torch

    g = golden("temporal_nt300.npz")
    gp = spb.StarryProcess(tau=0.7, **FID)
    gp._compute_moments()
    # the reference's own cov_ylm (see test_gpu_parity.py::test_sample_ylm_given_cov)
    gp._cov_ylm = torch.tensor(g["cov_ylm"], device="cuda").reshape(1, 256, 256).contiguous()
    gp._cho_cov_ylm = None
    y = gp.sample_ylm(t=g["ylm_t"], u=g["ylm_U"])
    assert tuple(y.shape) == g["ylm_y"].shape
    assert np.abs(y.cpu().numpy() - g["ylm_y"]).max() <= 1e-8 * np.abs(g["ylm_y"]).max()
    ys = gp.sample_ylm(t=g["ylm_t"], nsamples=5)
    assert tuple(ys.shape) == (5, 12, 256) and bool(torch.isfinite(ys).all())
