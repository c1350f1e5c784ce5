"""GPU-box probe: phase timers of moments_k1a (debug build libspb200_prof.so)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["SPB200_LIB"] = os.path.join(ROOT, "starry_process_b200", "libspb200_prof.so")
sys.path.insert(0, ROOT)
import numpy as np, torch
import starry_process_b200 as spb
B = 1184
rng = np.random.default_rng(7)
hp = dict(r=rng.uniform(10, 30, B), c=rng.uniform(0.01, 0.15, B), n=rng.uniform(1, 12, B),
          mu=rng.uniform(0, 85, B), sigma=rng.uniform(5, 40, B))
c = spb.get_context(0)
c.lib.spb_k1a_prof.argtypes = [ctypes.c_void_p]
out = (ctypes.c_ulonglong * 16)()
for rep in range(2):
    c.lib.spb_k1a_prof(out)
    gp = spb.StarryProcess(**hp)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); gp._compute_moments(); e1.record(); torch.cuda.synchronize()
    c.lib.spb_k1a_prof(out)
v = np.array(list(out), dtype=float)
names = ["profile+qs", "beta moments", "term table", "Y = Q Z", "S = Z^T Y + sym", "first moments + stores"]
print("moments total %.3f ms for B=%d" % (e0.elapsed_time(e1), B))
for k in range(6):
    print("  %-24s %9.1f kclk per sample  %5.1f%%" % (names[k], v[k] / B / 1e3, 100 * v[k] / v[:6].sum()))

print("K1b (one warp per sample): %.1f rounds per sample" % (v[14] / B))
for k, nm in enumerate(["rotation params", "row phase", "column phase", "loop overhead"]):
    print("  %-24s %9.1f kclk per sample  (%.0f clk per round)" % (nm, v[8 + k] / B / 1e3, v[8 + k] / max(v[14], 1)))
