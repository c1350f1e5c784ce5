"""GPU check of the equally-spaced-time-stamp fast path of the marginal assembly against the general kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import starry_process_b200 as spb
import bench
dev = torch.device("cuda")
for prior, seed in (("narrow", 1234), ("full", 4321)):
    B = 256
    hp, t, flux, _ = bench.synthetic_inputs(B, seed, prior)
    hd = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in hp.items()}
    res = {}
    for on in (0, 1):
        if on: os.environ.pop("SPB200_NO_UNIFORM_T", None)
        else: os.environ["SPB200_NO_UNIFORM_T"] = "1"
        gp = spb.StarryProcess(**hd)
        ll = gp.log_likelihood(t, flux, 1e-6, p=1.0, u=bench.U_LD)
        K = gp.cov(t[:300], p=1.0, u=bench.U_LD)
        torch.cuda.synchronize()
        res[on] = (ll.cpu().numpy(), K.cpu().numpy(), gp._udt)
    fin = np.isfinite(res[0][0])
    print("prior %-6s: uniform_dt %s / %s; lnlike max rel diff %.2e (finite %d, -inf pattern equal %s); cov max rel %.2e"
          % (prior, res[0][2], res[1][2], float(np.max(np.abs(res[1][0][fin] - res[0][0][fin]) / np.abs(res[0][0][fin]))),
             int(fin.sum()), bool(np.array_equal(np.isfinite(res[1][0]), fin)),
             float(np.nanmax(np.abs(res[1][1] - res[0][1]) / np.abs(res[0][1]).max()))), flush=True)
# different periods / non-integer ratios / irregular grid
hp, t, flux, _ = bench.synthetic_inputs(32, 1234, "narrow")
hd = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in hp.items()}
for p_, tt in ((0.37, t), (3.3, t), (1.0, np.linspace(-2.0, 7.5, 777)), (1.0, np.sort(np.random.default_rng(1).uniform(0, 4, 500)))):
    fl = np.interp(tt, t, flux)
    out = {}
    for on in (0, 1):
        if on: os.environ.pop("SPB200_NO_UNIFORM_T", None)
        else: os.environ["SPB200_NO_UNIFORM_T"] = "1"
        gp = spb.StarryProcess(**hd)
        out[on] = (gp.log_likelihood(tt, fl, 1e-6, p=p_, u=bench.U_LD).cpu().numpy(), gp._udt)
    fin = np.isfinite(out[0][0])
    print("p = %.2f nt = %d: uniform_dt %.6g; max rel diff %.2e" % (p_, len(tt), out[1][1], float(np.max(np.abs(out[1][0][fin] - out[0][0][fin]) / np.abs(out[0][0][fin])))), flush=True)
B = 4096
hp, t, flux, _ = bench.synthetic_inputs(B, 1234, "narrow")
hd = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in hp.items()}
td = torch.as_tensor(t, device=dev); fd = torch.as_tensor(flux, device=dev)
for on in (0, 1):
    if on: os.environ.pop("SPB200_NO_UNIFORM_T", None)
    else: os.environ["SPB200_NO_UNIFORM_T"] = "1"
    best = None
    for rep in range(4):
        gp = spb.StarryProcess(**hd); gp._stage_ms = {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        gp.log_likelihood(td, fd, 1e-6, p=1.0, u=bench.U_LD)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
        st = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in gp._stage_ms.items()}
        if best is None or ms < best[0]: best = (ms, st)
    print("uniform fast path %d: step %.2f ms, stages %s" % (on, best[0], {k: round(v, 2) for k, v in best[1].items()}), flush=True)
