"""GPU check of the INT8-tensor-core Cholesky path (spb_cholesky_lnlike_i8) against the FP64 (DMMA)
kernel: random SPD matrices through the C ABI, then the bench draws through StarryProcess."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import starry_process_b200 as spb  # noqa: E402
from starry_process_b200 import _lib  # noqa: E402
from starry_process_b200.sp import _ptr, _stream, get_context  # noqa: E402


def capi_case(ctx, B, n, M, planes, seed=0, dg=1e-4):
    lib, h = ctx.lib, ctx.handle
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(seed)
    ld = n + (n & 1)
    A = torch.randn(B, n, 24, dtype=torch.float64, generator=g)
    K0 = A @ A.transpose(1, 2) / 24.0
    K = torch.zeros(B, n, ld, dtype=torch.float64)
    K[:, :, :n] = K0
    r = torch.zeros(B, max(M, 1), ld, dtype=torch.float64)
    r[:, :, :n] = 0.01 * torch.randn(B, max(M, 1), n, dtype=torch.float64, generator=g)
    K, r = K.to(dev), r.to(dev)
    dgt = torch.full((1,), dg, dtype=torch.float64, device=dev)
    af = _lib.Affine()
    af.diag, af.diag_kind, af.diag_stride = dgt.data_ptr(), 0, 0
    out = {}
    for name in ("f64", "i8"):
        Kc, rc = K.clone(), r.clone()
        ll = torch.zeros(B, dtype=torch.float64, device=dev)
        quad = torch.zeros(B, max(M, 1), dtype=torch.float64, device=dev)
        logdet = torch.zeros(B, dtype=torch.float64, device=dev)
        info = torch.zeros(B, dtype=torch.int32, device=dev)
        if name == "f64":
            _lib.check(lib.spb_cholesky_lnlike_affine(h, B, n, _ptr(Kc), ld, n * ld, ctypes.byref(af), M,
                                                      _ptr(rc), ld, max(M, 1) * ld, _ptr(ll), _ptr(quad),
                                                      _ptr(logdet), _ptr(info), _stream()))
        else:
            nb = lib.spb_cholesky_i8_workspace_bytes(B, n, M, planes)
            ws = torch.empty(nb, dtype=torch.uint8, device=dev)
            _lib.check(lib.spb_cholesky_lnlike_i8(h, B, n, _ptr(Kc), ld, n * ld, ctypes.byref(af), M,
                                                  _ptr(rc), ld, max(M, 1) * ld, _ptr(ll), _ptr(quad),
                                                  _ptr(logdet), _ptr(info), planes, 0.0, _ptr(ws), nb, _stream()))
        torch.cuda.synchronize()
        out[name] = (ll.cpu(), quad.cpu(), logdet.cpu(), info.cpu(), rc.cpu())
    a, b = out["f64"], out["i8"]
    rel = lambda x, y: float(((x - y).abs() / x.abs().clamp_min(1e-300)).max())   # noqa: E731
    print("  B=%d n=%d M=%d planes=%d: lnlike rel %.2e  quad rel %.2e  logdet rel %.2e  y max abs %.2e  info %s / %s"
          % (B, n, M, planes, rel(a[0], b[0]) if M else 0.0, rel(a[1], b[1]) if M else 0.0, rel(a[2], b[2]),
             float((a[4][:, :M, :n] - b[4][:, :M, :n]).abs().max()) if M else 0.0,
             a[3].tolist()[:4], b[3].tolist()[:4]), flush=True)


def main():
    ctx = get_context()
    print("== C ABI, random SPD + 1e-4 I")
    for (B, n, M) in [(2, 128, 1), (2, 200, 1), (3, 257, 2), (2, 1000, 1), (2, 1000, 3), (1, 1024, 0),
                      (2, 1000, 130), (1, 2049, 1), (300, 320, 1)]:
        for planes in (8, 78, 7):
            capi_case(ctx, B, n, M, planes, seed=B * 1000 + n)
    print("== through StarryProcess (bench draws, marginalised + normalised)")
    import bench
    for prior, seed in (("narrow", 1234), ("full", 4321)):
        B = 256
        hp, t, flux, _ = bench.synthetic_inputs(B, seed, prior)
        dev = torch.device("cuda")
        hd = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in hp.items()}
        td = torch.as_tensor(t, dtype=torch.float64, device=dev)
        fd = torch.as_tensor(flux, dtype=torch.float64, device=dev)
        res = {}
        for planes in (0, 8, 78, 7):
            ctx.set_option("cholesky_i8", planes)
            ll = spb.StarryProcess(**hd).log_likelihood(td, fd, 1e-6, p=1.0, u=bench.U_LD)
            torch.cuda.synchronize()
            res[planes] = ll.cpu()
        fin = torch.isfinite(res[0])
        for planes in (8, 78, 7):
            same_inf = bool((torch.isfinite(res[planes]) == fin).all())
            rel = ((res[planes][fin] - res[0][fin]).abs() / res[0][fin].abs()).max()
            print("  prior %-6s planes %d: max rel lnlike diff vs FP64 kernel %.2e (finite %d / %d, -inf pattern equal: %s)"
                  % (prior, planes, float(rel), int(fin.sum()), B, same_inf), flush=True)
    print("== timing of the Cholesky stage inside log_likelihood (B = 1184 and 4096, nt = 1000)")
    for B in (1184, 4096):
        hp, t, flux, _ = bench.synthetic_inputs(B, 1234, "narrow")
        dev = torch.device("cuda")
        hd = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in hp.items()}
        td = torch.as_tensor(t, dtype=torch.float64, device=dev)
        fd = torch.as_tensor(flux, dtype=torch.float64, device=dev)
        for planes in (0, 8, 78, 7):
            ctx.set_option("cholesky_i8", planes)
            ts = []
            for rep in range(4):
                gp = spb.StarryProcess(**hd)
                gp._stage_ms = {}
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                gp.log_likelihood(td, fd, 1e-6, p=1.0, u=bench.U_LD)
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
                st = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in gp._stage_ms.items()}
            print("  B %d planes %d: step %.2f ms (best of 4), stages %s" % (B, planes, min(ts), {k: round(v, 2) for k, v in st.items()}), flush=True)
    ctx.set_option("cholesky_i8", -1)


if __name__ == "__main__":
    main()
