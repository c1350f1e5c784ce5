"""Single log_likelihood evaluation (configs[0]-like, marginal + conditional) for launch lists."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import starry_process_b200 as spb
g = np.load(os.path.join(ROOT, "tests", "golden", "fiducial_nt1000.npz"))
t = torch.tensor(g["t"], device="cuda"); f = torch.tensor(g["flux_norm"], device="cuda")
for rep in range(3):
    for marg in (True, False):
        gp = spb.StarryProcess(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0, marginalize_over_inclination=marg)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ll = gp.log_likelihood(t, f, 1e-6, i=60.0, p=1.0, u=[0.4, 0.26]); e1.record()
        torch.cuda.synchronize()
        print("marg=%d  %.3f ms  lnlike %.6f" % (marg, e0.elapsed_time(e1), ll.item()))
