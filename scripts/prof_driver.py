"""Short driver for ncu captures (run under `ncu ... python scripts/prof_driver.py [what]`).

what = sweep    : one chunk of the configs[2] sweep (B hyperparameter samples, nt=1000, marginal)
       design   : design matrix for nt=1e5 x I inclinations (configs[4])
       cond     : conditional lnlike, nt=1000 (A Sigma A^T + Cholesky)
Numbers printed by a run under ncu are never bench values.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import starry_process_b200 as spb  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "sweep"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 296
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
g = np.load(os.path.join(ROOT, "tests", "golden", "fiducial_nt1000.npz"))
rng = np.random.default_rng(7)
hp = dict(r=rng.uniform(10, 30, B), c=rng.uniform(0.01, 0.15, B), n=rng.uniform(1, 12, B),
          mu=rng.uniform(0, 85, B), sigma=rng.uniform(5, 40, B))
hp = {k: torch.tensor(v, device=dev) for k, v in hp.items()}
t = torch.tensor(g["t"], device=dev)
f = torch.tensor(g["flux_norm"], device=dev)

for _ in range(reps):
    if what == "sweep":
        gp = spb.StarryProcess(**hp)
        ll = gp.log_likelihood(t, f, 1e-6, p=1.0, u=[0.4, 0.26])
    elif what == "cond":
        gp = spb.StarryProcess(marginalize_over_inclination=False, **hp)
        ll = gp.log_likelihood(t, f, 1e-6, i=60.0, p=1.0, u=[0.4, 0.26])
    elif what == "design":
        gp = spb.StarryProcess(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0)
        inc = torch.rad2deg(torch.arccos(torch.rand(B, dtype=torch.float64, device=dev)))
        tt = torch.linspace(0, 40, 100000, dtype=torch.float64, device=dev)
        ll = gp.design_matrix(tt, i=inc, p=1.0, u=[0.4, 0.26])
    torch.cuda.synchronize()
print(what, "done", float(ll.reshape(-1)[0]))
