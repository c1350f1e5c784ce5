"""One unnormalised conditional log-likelihood call (B = 148, nt = 2048) for ncu: gemm_i8_lower_kernel, slice256_kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import starry_process_b200 as spb
import bench
hp, _, _, _ = bench.synthetic_inputs(148, 1234, "narrow")
nt = 2048
t = np.linspace(0, 8.0, nt)
f = 1e-3 * np.random.default_rng(5).standard_normal(nt)
for rep in range(2):
    gp = spb.StarryProcess(marginalize_over_inclination=False, normalized=False, **hp)
    ll = gp.log_likelihood(t, f, 1e-6, i=60.0, p=1.0, u=bench.U_LD)
torch.cuda.synchronize()
print(ll[:3].tolist())
