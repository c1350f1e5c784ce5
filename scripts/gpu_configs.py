"""GPU-box run of the remaining BASELINE.json configurations through the public StarryProcess API
(device-resident inputs, CUDA events, mean of `reps` after warm-up):
  configs[1]  1024 light curves sharing one hyperparameter set (one factorisation + 1024 RHS)
  configs[3]  512 hyperparameter samples x nt=4096, conditional i=60, quadratic limb darkening
  configs[0]  the reference's own single evaluation, for latency
Prints one JSON object per config."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import starry_process_b200 as spb

dev = torch.device("cuda:0")
U = [0.4, 0.26]
g = np.load(os.path.join(ROOT, "tests", "golden", "fiducial_nt1000.npz"))

def timed(fn, reps=3, warm=2):
    for _ in range(warm): out = fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out

# configs[0]: single evaluation latency (conditional, i = 60)
t = torch.tensor(g["t"], device=dev); f = torch.tensor(g["flux"], device=dev)
def one():
    gp = spb.StarryProcess(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0, marginalize_over_inclination=False, normalized=False)
    return gp.log_likelihood(t, f, 1e-6, i=60.0, p=1.0, u=U)
ms, ll = timed(one, reps=10)
print(json.dumps({"config": "configs[0] single lnlike nt=1000 conditional i=60", "ms_per_eval": ms,
                  "lnlike": float(ll), "golden": float(g["lnlike_m0_n0_uld"]),
                  "rel_err": abs(float(ll) - float(g["lnlike_m0_n0_uld"])) / abs(float(g["lnlike_m0_n0_uld"]))}), flush=True)

# configs[1]: ensemble
fens = torch.tensor(np.tile(g["flux_ens_norm"], (128, 1))[:1024], device=dev)
def ens():
    gp = spb.StarryProcess(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0)
    return gp.log_likelihood(t, fens, 1e-6, p=1.0, u=U)
ms, ll = timed(ens, reps=10)
print(json.dumps({"config": "configs[1] 1024 light curves, one K", "ms_per_step": ms, "evals_per_s": 1024 / ms * 1e3,
                  "lnlike_joint": float(ll)}), flush=True)

# configs[3]: long baseline
gl = np.load(os.path.join(ROOT, "tests", "golden", "longbaseline_nt4096.npz"))
t4 = torch.tensor(gl["t"], device=dev); f4 = torch.tensor(gl["flux"], device=dev)
B = 512
hp, _, _, _ = bench.synthetic_inputs(B, seed=99)
hp_d = {k: torch.tensor(v, device=dev) for k, v in hp.items()}
def long4096():
    gp = spb.StarryProcess(marginalize_over_inclination=False, normalized=False, **hp_d)
    return gp.log_likelihood(t4, f4, 1e-6, i=60.0, p=1.0, u=U)
ms, ll = timed(long4096, reps=2, warm=1)
nt = 4096
flop = B * (nt ** 3 / 3.0 + nt ** 2 + 2.0 * nt * 256 ** 2 + 1.0 * nt ** 2 * 256)
print(json.dumps({"config": "configs[3] 512 samples x nt=4096, conditional, LD", "ms_per_step": ms,
                  "evals_per_s": B / ms * 1e3, "algorithmic_TFLOPs": flop / ms / 1e9,
                  "finite": int(torch.isfinite(ll).sum())}), flush=True)
# parity of the three golden nt=4096 samples
gp = spb.StarryProcess(marginalize_over_inclination=False, normalized=False, r=gl["r"], mu=gl["mu"], sigma=gl["sigma"], c=gl["c"], n=gl["n"])
ll3 = gp.log_likelihood(t4, f4, 1e-6, i=60.0, p=1.0, u=U).cpu().numpy()
print(json.dumps({"config": "nt=4096 golden parity", "rel_err": float(np.max(np.abs(ll3 - gl["lnlike"]) / np.abs(gl["lnlike"])))}), flush=True)
