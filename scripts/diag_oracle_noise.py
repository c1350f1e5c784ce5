"""Development aid: how reproducible is the CPU oracle (== the reference's algorithm) on THIS host?
Prints lnlike under the reference's two eigensolver drivers, the clip probes and the noise-level
eigenvalues of the latitude moment matrix for one of the live-oracle test cases."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sp_oracle as so

rng = np.random.default_rng(11)
hp = dict(r=17.0, mu=42.0, sigma=11.0, c=0.08, n=6.0)
for nt in (1, 2, 63, 129):
    t = np.sort(rng.uniform(0, 3, nt)); f = 1e-3 * rng.standard_normal(nt)
    for marg in (False, True):
        for norm in (False, True):
            if nt == 1 and norm:
                continue
            if nt != 129:
                continue
            fn = lambda **kw: so.OracleProcess(marginalize_over_inclination=marg, normalized=norm, **hp, **kw).log_likelihood(t, f, 1e-6, i=33.0, p=0.7, u=[0.4, 0.26])
            base = fn()
            so.EIGH_DRIVER = "scipy"; alt = fn(); so.EIGH_DRIVER = "numpy"
            print("nt=%d marg=%d norm=%d numpy %.13f scipy %.13f floor %.2e" % (nt, marg, norm, base, alt, float(so.reference_noise_floor(fn))), flush=True)
o = so.OracleProcess(**hp)
w = np.linalg.eigvalsh(o.Q_lat)
print("eigvals[-31:-17] numpy:", w[-31:][:14])
import scipy.linalg
w2 = scipy.linalg.eigh(o.Q_lat, subset_by_index=(256 - 31, 255), eigvals_only=True)
print("eigvals scipy subset  :", w2[:14])
np.show_config()
