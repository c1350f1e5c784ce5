"""GPU check of the INT8-tensor-core SYRK of the Ylm moments (csrc/syrk_i8.cu) against the DMMA SYRK:
cov_ylm elementwise, lnlike of the bench draws, stage time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import starry_process_b200 as spb
import bench
ctx = spb.get_context()
dev = torch.device("cuda")
for prior, seed in (("narrow", 1234), ("full", 4321)):
    B = 256
    hp, t, flux, _ = bench.synthetic_inputs(B, seed, prior)
    hd = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in hp.items()}
    td = torch.as_tensor(t, device=dev); fd = torch.as_tensor(flux, device=dev)
    out = {}
    for on in (0, 1):
        ctx.set_option("moments_syrk_i8", on)
        gp = spb.StarryProcess(**hd)
        cov = gp.cov_ylm.clone()
        ll = gp.log_likelihood(td, fd, 1e-6, p=1.0, u=bench.U_LD)
        torch.cuda.synchronize()
        out[on] = (cov.cpu().numpy(), ll.cpu().numpy())
    c0, c1 = out[0][0], out[1][0]
    scale = np.abs(c0).max(axis=(1, 2), keepdims=True)
    fin = np.isfinite(out[0][1])
    print("prior %-6s: cov_ylm max |diff| / max|cov| = %.2e, symmetric %s; lnlike max rel diff %.2e (finite %d, -inf pattern equal %s)"
          % (prior, float((np.abs(c1 - c0) / scale).max()), bool(np.array_equal(c1, np.swapaxes(c1, 1, 2))),
             float(np.max(np.abs(out[1][1][fin] - out[0][1][fin]) / np.abs(out[0][1][fin]))), int(fin.sum()),
             bool(np.array_equal(np.isfinite(out[1][1]), fin))), flush=True)
# dr prior (ldeg factor)
hp, t, flux, _ = bench.synthetic_inputs(64, 1234, "narrow")
hd = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in hp.items()}
res = {}
for on in (0, 1):
    ctx.set_option("moments_syrk_i8", on)
    gp = spb.StarryProcess(dr=torch.full((64,), 4.0, dtype=torch.float64, device=dev), **hd)
    res[on] = gp.cov_ylm.cpu().numpy()
print("dr prior: cov_ylm max |diff| / max|cov| = %.2e" % float((np.abs(res[1] - res[0]) / np.abs(res[0]).max(axis=(1, 2), keepdims=True)).max()))
B = 4096
hp, t, flux, _ = bench.synthetic_inputs(B, 1234, "narrow")
hd = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in hp.items()}
td = torch.as_tensor(t, device=dev); fd = torch.as_tensor(flux, device=dev)
for on in (0, 1):
    ctx.set_option("moments_syrk_i8", on)
    best = None
    for rep in range(4):
        gp = spb.StarryProcess(**hd); gp._stage_ms = {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        gp.log_likelihood(td, fd, 1e-6, p=1.0, u=bench.U_LD)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
        st = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in gp._stage_ms.items()}
        if best is None or ms < best[0]: best = (ms, st)
    print("syrk_i8 %d: step %.2f ms, stages %s" % (on, best[0], {k: round(v, 2) for k, v in best[1].items()}), flush=True)
ctx.set_option("moments_syrk_i8", 0)
