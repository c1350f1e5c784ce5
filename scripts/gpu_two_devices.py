"""One process, two contexts (cuda:0 and cuda:1): same inputs must give identical results."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import starry_process_b200 as spb
g = np.load(os.path.join(ROOT, "tests", "golden", "fiducial_nt1000.npz"))
out = []
for dev in (0, 1, 0):
    gp = spb.StarryProcess(r=[10.0, 15.0], mu=[30.0, 40.0], sigma=[5.0, 8.0], c=[0.1, 0.05], n=[10.0, 4.0],
                           device=dev)
    ll = gp.log_likelihood(g["t"], g["flux_norm"], 1e-6, p=1.0, u=[0.4, 0.26])
    A = gp.design_matrix(g["t"][:50], i=60.0)
    mu, K = spb.StarryProcess(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0, normalized=False, device=dev).predict(
        g["t"][:300], g["flux"][:300], 1e-6, t_sample=g["t"][:40])
    out.append((ll.cpu().numpy(), float(A.sum()), float(mu.sum()), str(ll.device)))
    print(out[-1])
assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][0], out[2][0])
assert out[0][1] == out[1][1] and out[0][2] == out[1][2]
print("two-device check ok")
