"""CPU, build container only (skipped where /root/reference or oracle/_ref are absent): the plain-C
restatement against the reference's own compiled headers, and the NumPy restatement against the
unmodified reference package executed through oracle/theano_stub."""
import os

import numpy as np
import pytest

from conftest import FID, REFERENCE, U_LD


@pytest.fixture(scope="module")
def refnat(oracle):
    if not oracle.ref_available(15, 2):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle.get_native("ref")


def test_native_restatement_bitwise(oracle, refnat):
    nc = oracle.get_native("port")
    for th in (0.5 * np.pi, -np.pi / 3, 0.1, 0.0, 2.5):
        assert np.array_equal(nc.Rx(15, 2, th), refnat.Rx(15, 2, th))
    for al, be in ((53.6, 7.2), (1.0, 0.5), (22026.0, 22026.0), (3.3, 0.9), (1.0000001, 0.50001)):
        q1, Q1 = nc.latitude(15, 2, al, be)
        q2, Q2 = refnat.latitude(15, 2, al, be)
        assert np.array_equal(q1, q2) and np.array_equal(Q1, Q2)
    for args in ((-0.5, 7.2, 60.8, 0.5), (-0.5, 0.5, 1.5, 0.5), (-0.5, 300.0, 301.0, 0.5)):
        assert nc.hyp2f1(*args) == refnat.hyp2f1(*args)
    rng = np.random.default_rng(1)
    M = rng.standard_normal((50, 256))
    th = rng.uniform(0, 2 * np.pi, 50)
    assert np.array_equal(nc.tensordotRz(15, 2, M, th), refnat.tensordotRz(15, 2, M, th))
    T_ = rng.standard_normal((256, 256))
    M2 = rng.standard_normal((256, 256))
    f1 = nc.special_tensordotRz(15, 2, T_, M2, th)
    f2 = refnat.special_tensordotRz(15, 2, T_, M2, th)
    assert np.abs(f1 - f2).max() <= 1e-12 * np.abs(f2).max()
    assert np.abs(nc.rTA1(15, 2) - refnat.rTA1(15, 2)).max() <= 1e-13
    for u in ([0.0, 0.0], [0.4, 0.26], [1.0, -0.3]):
        assert np.abs(nc.rTA1L(15, 2, u) - refnat.rTA1L(15, 2, u)).max() <= 2e-13


def test_small_degree_restatement(oracle):
    if not oracle.ref_available(5, 2):
        pytest.skip("oracle/_ref (ydeg=5) not built")
    nc, nr = oracle.get_native("port"), oracle.get_native("ref")
    assert np.array_equal(nc.Rx(5, 2, 0.7), nr.Rx(5, 2, 0.7))
    q1, Q1 = nc.latitude(5, 2, 12.0, 3.0)
    q2, Q2 = nr.latitude(5, 2, 12.0, 3.0)
    assert np.array_equal(Q1, Q2) and np.array_equal(q1, q2)
    assert np.abs(nc.rTA1L(5, 2, [0.3, 0.1]) - nr.rTA1L(5, 2, [0.3, 0.1])).max() <= 1e-14


def test_live_reference_through_stub(oracle, have_reference):
    if not have_reference or not oracle.ref_available(15, 2):
        pytest.skip("reference tree not available")
    from oracle import theano_stub

    sp = theano_stub.import_reference()
    t = np.linspace(0, 4, 300)
    flux = np.random.default_rng(0).standard_normal(300) * 1e-3
    hp = dict(r=23.0, mu=51.0, sigma=9.0, c=0.07, n=4.0)
    for marg in (False, True):
        for norm in (False, True):
            kw = dict(marginalize_over_inclination=marg, normalized=norm, **hp)
            g = sp.StarryProcess(**kw)
            o = oracle.OracleProcess(**kw)
            if not marg and not norm:
                assert np.array_equal(np.array(g.mean_ylm.eval()), o.mean_ylm)
                assert np.array_equal(np.array(g.cov_ylm.eval()), o.cov_ylm)
            l1 = float(g.log_likelihood(t, flux, 1e-6, i=40.0, p=1.3, u=U_LD))
            l2 = o.log_likelihood(t, flux, 1e-6, i=40.0, p=1.3, u=U_LD)
            assert abs(l1 - l2) <= 1e-12 * abs(l1)
