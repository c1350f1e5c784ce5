#!/usr/bin/env python
"""Regenerates starry_process_b200/data/longitude_U_ydeg15.npy, the pinned longitude eigenvector
table ``U_lon = matrix_sqrt(Q_lon)`` (longitude.py:9-49 -> integrals.py:116-124 -> math.py:121-139,
numpy.linalg.eigh driver of ops/eigh/eigh.py:11-20), and records its provenance in the ``.json`` next
to it (NumPy / OpenBLAS versions, BLAS thread count, CPU, SHA-256 of the table).

    python scripts/gen_longitude_U.py            # writes the table with the container's default BLAS threads
    python scripts/gen_longitude_U.py --check    # exit 0 iff this host reproduces the shipped table bitwise

``Q_lon`` (256 x 256, constant) has rank 31 with several eigenvalues at the 1e-15 clip; the
eigenvectors LAPACK returns for them depend on the BLAS kernels and thread count (the projector
U U^T is reproducible to 1e-15, U itself only to 0.2), and the reference's log-likelihood moves by up
to 3e-6 with them.  Every golden fixture under tests/golden/ was produced in the build container
with the default thread count (8); this script run there, the same way, reproduces the committed
table bit for bit (``--check``).  ``StarryProcess(longitude_basis="host")`` bypasses the table and
uses this host's own eigh instead.
"""
import hashlib
import json
import os
import platform
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from starry_process_b200 import _tables as T  # noqa: E402


def provenance(U):
    info = {"sha256": hashlib.sha256(np.ascontiguousarray(U).tobytes()).hexdigest(),
            "shape": list(U.shape), "numpy": np.__version__, "python": platform.python_version(),
            "machine": platform.machine(), "processor": platform.processor()}
    try:
        import scipy

        info["scipy"] = scipy.__version__
    except ImportError:
        pass
    try:
        from threadpoolctl import threadpool_info

        info["blas"] = [{k: d.get(k) for k in ("internal_api", "version", "num_threads",
                                               "threading_layer", "architecture", "filepath")}
                        for d in threadpool_info() if d.get("user_api") == "blas"]
        for d in info["blas"]:
            d["filepath"] = os.path.basename(d["filepath"] or "")
    except ImportError:
        info["blas"] = "threadpoolctl unavailable"
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    info["cpu"] = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    info["env_threads"] = {k: os.environ.get(k) for k in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS",
                                                          "MKL_NUM_THREADS")}
    return info


def main():
    U = T.longitude_U("host")
    if "--check" in sys.argv:
        Up = T.longitude_U("pinned")
        same = np.array_equal(U, Up)
        print("this host %s the pinned table bitwise; max |dU| = %.3e, max |d(U U^T)| = %.3e"
              % ("reproduces" if same else "does NOT reproduce", np.abs(U - Up).max(),
                 np.abs(U @ U.T - Up @ Up.T).max()))
        return 0 if same else 1
    np.save(T.PINNED_LONGITUDE, U)
    with open(T.PINNED_LONGITUDE[:-4] + ".json", "w") as fh:
        json.dump(provenance(U), fh, indent=1, sort_keys=True)
        fh.write("\n")
    print("wrote", T.PINNED_LONGITUDE)
    return 0


if __name__ == "__main__":
    sys.exit(main())
