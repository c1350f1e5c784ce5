"""CPU: the gradient oracle (oracle/sp_oracle_grad.py) is pinned the way the reference pins its own
gradient -- against finite differences of the forward function (theano.gradient.verify_grad,
tests/test_lnlike.py:105-136: tolerance 1e-4) -- here central differences of the UNMODIFIED
reference's log_likelihood run through oracle/theano_stub, in all four marginalise x normalise modes."""
import numpy as np
import pytest

from conftest import U_LD

HP = dict(r=20.0, a=0.40, b=0.27, c=0.1, n=10.0)     # the reference's defaults (defaults.py:4-35)
STEPS = dict(r=1e-3, a=1e-5, b=1e-5, c=1e-6, n=1e-4)
VERIFY_GRAD_TOL = 1e-4


def _data():
    rng = np.random.default_rng(42)
    t = np.linspace(0, 3, 100)            # the reference's test grid
    return t, 1e-3 * rng.standard_normal(100)


@pytest.mark.parametrize("marg", [False, True])
@pytest.mark.parametrize("norm", [False, True])
def test_gradient_oracle_vs_reference_finite_differences(oracle, have_reference, marg, norm):
    from oracle import sp_oracle_grad as sg

    t, flux = _data()
    ll, g = sg.lnlike_and_grad(HP, t, flux, 1e-6, i=60.0, p=1.0, u=U_LD,
                               marginalize_over_inclination=marg, normalized=norm)
    if have_reference:
        from oracle import theano_stub

        SP = theano_stub.import_reference().StarryProcess

        def fwd(**hp):
            return float(SP(ydeg=15, marginalize_over_inclination=marg, normalized=norm,
                            **hp).log_likelihood(t, flux, 1e-6, i=60.0, p=1.0, u=U_LD))
    else:
        def fwd(**hp):
            return oracle.OracleProcess(marginalize_over_inclination=marg, normalized=norm,
                                        **hp).log_likelihood(t, flux, 1e-6, i=60.0, p=1.0, u=U_LD)
    assert abs(ll - fwd(**HP)) <= 1e-8 * abs(ll)
    for p_ in sg.PARAMS:
        hp1, hp2 = dict(HP), dict(HP)
        hp1[p_] += STEPS[p_]
        hp2[p_] -= STEPS[p_]
        fd = (fwd(**hp1) - fwd(**hp2)) / (2 * STEPS[p_])
        assert abs(g[p_] - fd) <= VERIFY_GRAD_TOL * max(abs(fd), 1.0), (p_, g[p_], fd)


def test_latitude_derivative_lanes_vs_reference_cpp(oracle):
    """dQ/dalpha, dQ/dbeta of the oracle (even/even lanes of latitude.h:48-60, 112-172) against
    central differences of the reference's compiled computeLatitudeIntegrals (oracle/_ref)."""
    from oracle import sp_oracle_grad as sg

    nat = oracle.get_native("ref" if oracle.ref_available(15, 2) else "port")
    al, be = 53.59815003, 3.22602246
    q3, Q3 = sg.latitude_with_grad(15, al, be)
    q0, Q0 = nat.latitude(15, 2, al, be)
    low = slice(0, 100)
    assert np.abs(Q3[0] - Q0)[low, low].max() <= 1e-13 * np.abs(Q0).max()
    for lane, (da, db) in ((1, (1e-4, 0.0)), (2, (0.0, 1e-5))):
        qp, Qp = nat.latitude(15, 2, al + da, be + db)
        qm, Qm = nat.latitude(15, 2, al - da, be - db)
        h = da + db
        fdQ = (Qp - Qm) / (2 * h)
        assert np.abs(Q3[lane] - fdQ)[low, low].max() <= 1e-6 * np.abs(fdQ[low, low]).max()
        assert np.abs(q3[lane] - (qp - qm) / (2 * h))[low].max() <= 1e-6 * np.abs(q3[lane][low]).max()
