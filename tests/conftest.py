import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("SP_REFERENCE_ROOT", "/root/reference")


# The golden fixtures were produced by the unmodified reference with the BLAS thread count of the
# build container (8).  The reference's marginalised lnlike depends on that count at the 3e-6 level
# (different LAPACK blocking -> different noise-level eigenmodes; DESIGN.md "numerical fragility"),
# so the CPU comparisons of the oracle against the fixtures pin it -- whatever OMP_NUM_THREADS /
# OPENBLAS_NUM_THREADS the caller's environment carries.
GOLDEN_BLAS_THREADS = 8
_BLAS_LIMIT = None


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    global _BLAS_LIMIT
    try:
        import numpy  # noqa: F401  (load every BLAS copy first: threadpoolctl only limits what is
        import scipy.linalg  # noqa: F401   already loaded; NumPy and SciPy ship their own OpenBLAS)
        from threadpoolctl import threadpool_limits

        _BLAS_LIMIT = threadpool_limits(limits=GOLDEN_BLAS_THREADS, user_api="blas")
    except Exception:  # threadpoolctl missing: the environment's default applies
        _BLAS_LIMIT = None


def _ensure_oracle_native():
    so_path = os.path.join(ROOT, "oracle", "liboracle_native.so")
    src = os.path.join(ROOT, "oracle", "oracle_native.c")
    if not os.path.exists(so_path) or os.path.getmtime(so_path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "native"],
                              stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle():
    _ensure_oracle_native()
    from oracle import sp_oracle

    return sp_oracle


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))

    return load


@pytest.fixture(scope="session")
def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "starry_process"))


FID = dict(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0)
U_LD = [0.4, 0.26]
