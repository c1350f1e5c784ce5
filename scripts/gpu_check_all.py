"""GPU-box development check: every libspb200 stage against the CPU oracle (prints, no asserts)."""
import os, sys, time, ctypes
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import starry_process_b200 as spb
from starry_process_b200 import _lib, _tables
from oracle import sp_oracle as so

torch.set_printoptions(precision=12)
dev = torch.device("cuda:0")
G = lambda name: np.load(os.path.join(ROOT, "tests", "golden", name))
fid = G("fiducial_nt1000.npz")
t = fid["t"]; 
FID = dict(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0)
P = lambda x: ctypes.c_void_p(x.data_ptr())

def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)

ctx = spb.get_context(0)
lib, h = ctx.lib, ctx.handle
nat = so.get_native("port")

# ---- Rx
th = torch.tensor([0.5 * np.pi, -np.pi / 3, 0.1, 0.0, 1.2345], dtype=torch.float64, device=dev)
Rx = torch.empty(5, 5456, dtype=torch.float64, device=dev)
_lib.check(lib.spb_Rx(h, 5, P(th), P(Rx), None)); torch.cuda.synchronize()
for k in range(5):
    print("Rx(%.4f) maxdiff %.2e" % (th[k].item(), np.abs(Rx[k].cpu().numpy() - nat.Rx(15, 2, th[k].item())).max()))
# ---- flux operator
for u in ([0.0, 0.0], [0.4, 0.26]):
    ut = torch.tensor(u, dtype=torch.float64, device=dev); out = torch.empty(256, dtype=torch.float64, device=dev)
    _lib.check(lib.spb_flux_operator(h, 1, P(ut), P(out), None)); torch.cuda.synchronize()
    print("rTA1L", u, "maxdiff %.2e" % np.abs(out.cpu().numpy() - nat.rTA1L(15, 2, u)).max())
# ---- tensordotRz
rng = np.random.default_rng(3)
M = rng.standard_normal((40, 256)); thz = rng.uniform(0, 2 * np.pi, 40)
f = torch.empty(40, 256, dtype=torch.float64, device=dev)
_lib.check(lib.spb_tensordotRz(h, 40, P(torch.tensor(M, device=dev)), P(torch.tensor(thz, device=dev)), P(f), None)); torch.cuda.synchronize()
print("tensordotRz maxdiff %.2e" % np.abs(f.cpu().numpy() - nat.tensordotRz(15, 2, M, thz)).max())

# ---- moments, fiducial
t0 = time.time()
gp = spb.StarryProcess(marginalize_over_inclination=False, normalized=False, **FID)
mu = gp.mean_ylm.cpu().numpy(); cov = gp.cov_ylm.cpu().numpy(); torch.cuda.synchronize()
print("moments fiducial: %.2fs  mean maxdiff %.2e (scale %.2e)  cov maxdiff %.2e (scale %.2e) info %s" % (
    time.time() - t0, np.abs(mu - fid["mean_ylm"]).max(), np.abs(mu).max(), np.abs(cov - fid["cov_ylm"]).max(), np.abs(cov).max(), gp.info.cpu().numpy()))
# ---- design matrix
dm = G("design_matrix_ref.npz")
for k, u in enumerate(([0.0, 0.0], dm["u"])):
    A = gp.design_matrix(dm["t"], i=torch.tensor(dm["incs"]), p=1.0, u=u).cpu().numpy()
    print("design matrix u=%s maxdiff vs reference %.2e" % (list(u), np.abs(A - dm["A"][k]).max()))
af = G("design_matrix_AF15.npz")
A = gp.design_matrix(af["theta_deg"] / 360.0, i=torch.tensor(af["incs"]), p=1.0, u=[0.0, 0.0]).cpu().numpy()
print("design matrix vs starry A_F15 fixture maxdiff %.2e" % np.abs(A - af["A_F"]).max())
# ---- cov / lnlike all combos
for marg in (False, True):
    for norm in (False, True):
        g = spb.StarryProcess(marginalize_over_inclination=marg, normalized=norm, **FID)
        for uname, u in (("u0", [0.0, 0.0]), ("uld", [0.4, 0.26])):
            key = "m%d_n%d_%s" % (marg, norm, uname)
            K = g.cov(t, i=60.0, p=1.0, u=u).cpu().numpy()
            d0 = np.abs(K[0] - fid["Krow0_" + key]).max(); d5 = np.abs(K[500] - fid["Krow500_" + key]).max()
            dd = np.abs(np.diag(K) - fid["Kdiag_" + key]).max()
            fl = fid["flux_norm"] if norm else fid["flux"]; fe = fid["flux_ens_norm"] if norm else fid["flux_ens"]
            ll = g.log_likelihood(t, fl, 1e-6, i=60.0, p=1.0, u=u).item()
            lle = g.log_likelihood(t, fe, 1e-6, i=60.0, p=1.0, u=u).item()
            print("%s: K rows maxdiff %.2e %.2e diag %.2e (scale %.2e) | lnlike %.10f ref %.10f rel %.2e | ens rel %.2e" % (
                key, d0, d5, dd, np.abs(K).max(), ll, fid["lnlike_" + key], rel(ll, fid["lnlike_" + key]), rel(lle, fid["lnlike_ens_" + key])))
        key = "m%d_n%d" % (marg, norm)
        fl = fid["flux_norm"] if norm else fid["flux"]
        ll = g.log_likelihood(t, fl, fid["data_cov_vec"], i=60.0, p=1.0, u=[0.4, 0.26], baseline_mean=1e-4, baseline_var=1e-5).item()
        print("   dvec/baseline lnlike rel %.2e" % rel(ll, fid["lnlike_dvec_" + key]))
# ---- sweeps (batched)
for name in ("sweep_nt1000.npz", "sweep_lowc_nt1000.npz"):
    sw = G(name)
    for marg in (False, True):
        for norm in (False, True):
            g = spb.StarryProcess(r=sw["r"], mu=sw["mu"], sigma=sw["sigma"], c=sw["c"], n=sw["n"], marginalize_over_inclination=marg, normalized=norm)
            fl = fid["flux_norm"] if norm else fid["flux"]
            ll = g.log_likelihood(t, fl, 1e-6, i=60.0, p=1.0, u=[0.4, 0.26]).cpu().numpy()
            ref = sw["lnlike_m%d_n%d" % (marg, norm)]
            fin = np.isfinite(ref)
            r_ = np.abs(ll[fin] - ref[fin]) / np.abs(ref[fin])
            print("%s m%d n%d: finite %d/%d inf-match %s | rel err max %.2e median %.2e | info %s" % (
                name, marg, norm, fin.sum(), len(ref), bool(np.all(np.isinf(ll[~fin]))), r_.max() if fin.any() else 0, np.median(r_) if fin.any() else 0, np.unique(g.info.cpu().numpy())))
    g = spb.StarryProcess(r=sw["r"], mu=sw["mu"], sigma=sw["sigma"], c=sw["c"], n=sw["n"])
    mu_ = g.mean_ylm.cpu().numpy(); cv = g.cov_ylm.cpu().numpy()
    print("   mean_ylm rel %.2e ; cov diag rel-to-max %.2e ; row6 rel-to-max %.2e" % (
        np.abs(mu_ - sw["mean_ylm"]).max() / np.abs(sw["mean_ylm"]).max(),
        (np.abs(np.diagonal(cv, axis1=1, axis2=2) - sw["cov_ylm_diag"]).max(1) / np.abs(sw["cov_ylm_diag"]).max(1)).max(),
        (np.abs(cv[:, 6, :] - sw["cov_ylm_row6"]).max(1) / np.abs(sw["cov_ylm_row6"]).max(1)).max()))
# ---- sample_ylm
sy = G("sample_ylm.npz")
gp = spb.StarryProcess(**FID)
y = gp.sample_ylm(u=sy["unit_normals"]).cpu().numpy()
print("sample_ylm maxdiff %.2e (scale %.2e)" % (np.abs(y - sy["y"]).max(), np.abs(sy["y"]).max()))
# ---- long baseline
lb = G("longbaseline_nt4096.npz")
g = spb.StarryProcess(r=lb["r"], mu=lb["mu"], sigma=lb["sigma"], c=lb["c"], n=lb["n"], marginalize_over_inclination=False, normalized=False)
ll = g.log_likelihood(lb["t"], lb["flux"], 1e-6, i=60.0, p=1.0, u=lb["u"]).cpu().numpy()
print("nt=4096 lnlike rel err", np.abs(ll - lb["lnlike"]) / np.abs(lb["lnlike"]))
# ---- timing of the stages, B = 1024
B = 1024
rng = np.random.default_rng(0)
hp = dict(r=rng.uniform(10, 30, B), c=rng.uniform(0.01, 0.15, B), n=rng.uniform(1, 12, B), mu=rng.uniform(0, 85, B), sigma=rng.uniform(5, 40, B))
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    g = spb.StarryProcess(**hp); g._compute_moments(); torch.cuda.synchronize(); t1 = time.time()
    ll = g.log_likelihood(t, fid["flux_norm"], 1e-6, p=1.0, u=[0.4, 0.26]); torch.cuda.synchronize(); t2 = time.time()
    print("B=%d marginal+normalized: moments %.1f ms, lnlike %.1f ms -> %.0f evals/s ; finite %d" % (B, (t1 - t0) * 1e3, (t2 - t1) * 1e3, B / (t2 - t0), int(torch.isfinite(ll).sum())))
print("kernel launches:", ctx.launches())
